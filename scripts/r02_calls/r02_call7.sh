#!/bin/bash
# round-2 GPU call 7 (2 GPUs): multi-GPU tests (torchrun K7/NCCL + in-process sharded index over 2 devices), bench at N=2
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_n2.txt 2>&1
tail -15 gpurun_out/r02_pytest_n2.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -8 gpurun_out/r02_bench_n2.err
