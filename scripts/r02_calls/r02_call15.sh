#!/bin/bash
# round-2 GPU call 15 (8 GPUs): bench at N=8 as the driver launches it, C5 (100M x 768 over 8 GPUs), sharded + multi tests
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n8b.json 2> gpurun_out/r02_bench_n8b.err
tail -4 gpurun_out/r02_bench_n8b.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload c5 --steps 10 --warmup 3 --no-cpu ) > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err
tail -4 gpurun_out/r02_bench_c5_n8.err
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_n8b.txt 2>&1
tail -5 gpurun_out/r02_pytest_n8b.txt
