#!/bin/bash
# round-2 GPU call 21 (8 GPUs): default bench at N=8 after the 16 KiB K7 instance / K2 on every SM; multi + sharded tests
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 8 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n8c.json 2> gpurun_out/r02_bench_n8c.err
tail -4 gpurun_out/r02_bench_n8c.err
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_n8c.txt 2>&1
tail -4 gpurun_out/r02_pytest_n8c.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29526 bench.py --gpus 8 --workload t10mb --steps 50 --warmup 5 --no-cpu --no-parity ) > gpurun_out/r02_bench_t10mb_n8.json 2> gpurun_out/r02_bench_t10mb_n8.err
tail -3 gpurun_out/r02_bench_t10mb_n8.err
