#!/bin/bash
# round-2 GPU call 14 (2 GPUs): whole GPU suite on a 2-GPU box, bench at N=2 (driver-style launch), K8 cache on where-based C4b
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r02_pytest9.txt 2>&1
tail -12 gpurun_out/r02_pytest9.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 100 --warmup 10 ) > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err
tail -4 gpurun_out/r02_bench_n2b.err
timeout 600 python scripts/k8_cache.py > gpurun_out/r02_k8_cache.txt 2>&1
cat gpurun_out/r02_k8_cache.txt
