#!/bin/bash
# round-2 GPU call 17 (2 GPUs): t10mb / default bench after the 16 KiB K7 instance; multi test
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_n2c.txt 2>&1
tail -3 gpurun_out/r02_pytest_n2c.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 2 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n2c.json 2> gpurun_out/r02_bench_n2c.err
tail -3 gpurun_out/r02_bench_n2c.err
