#!/bin/bash
# round-2 GPU call 26 (8 GPUs): C5 (100M x 768, 1024-query batches) with the multi-chunk pair kernel; default bench at N=8
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --workload c5 --steps 10 --warmup 3 --no-cpu ) > gpurun_out/r02_bench_c5_n8b.json 2> gpurun_out/r02_bench_c5_n8b.err
tail -3 gpurun_out/r02_bench_c5_n8b.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n8d.json 2> gpurun_out/r02_bench_n8d.err
tail -3 gpurun_out/r02_bench_n8d.err
