#!/bin/bash
# round-2 GPU call 28 (1 GPU): default bench with the 1024-query leg in `also`
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 900 python bench.py --gpus 1 --steps 50 --warmup 8 ) > gpurun_out/r02_bench_final3.json 2> gpurun_out/r02_bench_final3.err
tail -4 gpurun_out/r02_bench_final3.err
