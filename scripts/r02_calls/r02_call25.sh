#!/bin/bash
# round-2 GPU call 25 (1 GPU): multi-chunk pair launches (up to 1024 queries per corpus pass) — parity, then C5-shard / C3 timing
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_config_shapes.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_mc.txt 2>&1
tail -15 gpurun_out/r02_pytest_mc.txt
timeout 600 python bench.py --workload c5s --steps 8 --warmup 3 --no-cpu > gpurun_out/r02_c5s_mc.json 2> gpurun_out/r02_c5s_mc.err
tail -2 gpurun_out/r02_c5s_mc.err
YRB_K2_MULTICHUNK=0 timeout 600 python bench.py --workload c5s --steps 8 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_c5s_sc.json 2> gpurun_out/r02_c5s_sc.err
timeout 300 python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu --no-parity > gpurun_out/r02_c3_mc.json 2> gpurun_out/r02_c3_mc.err
