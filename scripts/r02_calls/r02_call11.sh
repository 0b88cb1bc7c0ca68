#!/bin/bash
# round-2 GPU call 11: whole GPU suite (new: config shapes, 1000 collections, K5 stream, K1 small grids, pooled stream), crossover, ingest rate
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02_pytest6.txt 2>&1
tail -25 gpurun_out/r02_pytest6.txt
timeout 300 python scripts/crossover.py 1000 10000 > gpurun_out/r02_crossover3.txt 2>&1
cat gpurun_out/r02_crossover3.txt
timeout 300 python bench.py --workload c2 --steps 50 --warmup 5 --no-cpu --no-parity > gpurun_out/r02_c2.json 2> gpurun_out/r02_c2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c2_launches.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_c2.log 2>&1
grep ingest gpurun_out/r02_c2_launches.csv | tail -3
