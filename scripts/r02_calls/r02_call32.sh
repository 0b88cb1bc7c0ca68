#!/bin/bash
# round-2 GPU call 32 (2 GPUs): the default line under torchrun with the t10mq side leg (1024-query batches through K7)
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
echo rc=$?
tail -3 gpurun_out/r02_bench_n2_final.err
