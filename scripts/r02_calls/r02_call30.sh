#!/bin/bash
# round-2 GPU call 30 (1 GPU): ncu --set full of the multi-chunk pair kernel (10M x 1024, 1024 queries in one launch):
# dram bytes of the launch against the 20.48 GB of rows
set -x
cd "$GRAFT_REPO_ROOT"
timeout 170 ncu --set full --clock-control none --import-source on -k regex:k2_gemm_topk_pair -s 3 -c 1 -o gpurun_out/r02_k2pair_t10mq -f python bench.py --workload t10mq --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_k2mc.log 2>&1
echo rc=$?
ls -la gpurun_out/
