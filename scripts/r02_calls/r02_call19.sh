#!/bin/bash
# round-2 GPU call 19 (1 GPU): smoke and a small batch UNDER ncu (one-CTA K2 with compile-time cluster dims), then the suite
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_ncu_smoke.log 2>&1
echo "ncu smoke rc=$?"; tail -2 gpurun_out/r02_ncu_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small16_launches.csv python scripts/small_batch.py 10000 16 > gpurun_out/r02_ncu_small16.log 2>&1
echo "ncu small16 rc=$?"; tail -2 gpurun_out/r02_ncu_small16.log
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest_final2.txt 2>&1
tail -5 gpurun_out/r02_pytest_final2.txt
timeout 300 python scripts/crossover.py 1000 10000 > gpurun_out/r02_crossover5.txt 2>&1
tail -10 gpurun_out/r02_crossover5.txt
