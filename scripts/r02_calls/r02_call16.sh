#!/bin/bash
# round-2 GPU call 16 (4 GPUs): bench at N=4 as the driver launches it
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err
tail -4 gpurun_out/r02_bench_n4.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 4 --steps 5 --warmup 3 ) > gpurun_out/r02_bench_ref_n4.json 2> gpurun_out/r02_bench_ref_n4.err
tail -3 gpurun_out/r02_bench_ref_n4.err
