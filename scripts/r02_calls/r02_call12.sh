#!/bin/bash
# round-2 GPU call 12: suite after the filter caches / K5 reciprocal / merged config-shape probes; ingest + small-batch timings
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r02_pytest7.txt 2>&1
tail -14 gpurun_out/r02_pytest7.txt
timeout 300 python scripts/crossover.py 1000 10000 > gpurun_out/r02_crossover4.txt 2>&1
cat gpurun_out/r02_crossover4.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c2_launches2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_c2b.log 2>&1
grep ingest gpurun_out/r02_c2_launches2.csv | tail -2
