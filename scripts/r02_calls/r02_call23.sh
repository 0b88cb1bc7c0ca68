#!/bin/bash
# round-2 GPU call 23 (1 GPU): final state — suite, smoke, default bench (driver-like flags), C2 for continuity with round 1
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest_final4.txt 2>&1
tail -4 gpurun_out/r02_pytest_final4.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final2.txt 2>&1; tail -1 gpurun_out/r02_smoke_final2.txt
timeout 900 python bench.py --gpus 1 --steps 100 --warmup 10 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -3 gpurun_out/r02_bench_final.err
timeout 300 python bench.py --workload c2 --steps 300 --warmup 30 --no-parity > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err
