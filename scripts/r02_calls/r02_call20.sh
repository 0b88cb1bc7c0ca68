#!/bin/bash
# round-2 GPU call 20 (1 GPU): suite + crossover + smoke under ncu after the K3 small-n path; default bench line
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest_final3.txt 2>&1
tail -5 gpurun_out/r02_pytest_final3.txt
timeout 300 python scripts/crossover.py 1000 10000 100000 > gpurun_out/r02_crossover6.txt 2>&1
cat gpurun_out/r02_crossover6.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_ncu_smoke.log 2>&1
echo "ncu smoke rc=$?"
timeout 1200 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_default3.json 2> gpurun_out/r02_bench_default3.err
tail -3 gpurun_out/r02_bench_default3.err
