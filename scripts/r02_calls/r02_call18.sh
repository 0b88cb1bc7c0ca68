#!/bin/bash
# round-2 GPU call 18 (1 GPU): what the driver runs at round end — pytest -m gpu, smoke(), bench with a short step count, the reference arm
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest_final.txt 2>&1
tail -6 gpurun_out/r02_pytest_final.txt
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_smoke_final.txt 2>&1
tail -5 gpurun_out/r02_smoke_final.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_driverlike.json 2> gpurun_out/r02_bench_driverlike.err
tail -4 gpurun_out/r02_bench_driverlike.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_ref_driverlike.json 2> gpurun_out/r02_bench_ref_driverlike.err
tail -4 gpurun_out/r02_bench_ref_driverlike.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_ncu_smoke.log 2>&1
