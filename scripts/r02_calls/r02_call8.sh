#!/bin/bash
# round-2 GPU call 8 (8 GPUs): bench at N=8 as the driver launches it, in-process sharded index over 8 devices, torchrun parity test
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L | wc -l
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -5 gpurun_out/r02_bench_n8.err
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_n8.txt 2>&1
tail -15 gpurun_out/r02_pytest_n8.txt
