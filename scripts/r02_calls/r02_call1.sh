#!/bin/bash
# round-2 GPU call 1: UMMA/shared-memory probe, baseline GPU tests, K2 one-CTA vs pair, ncu of the pair kernel
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_box.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I youtu-rag_b200/csrc -o /tmp/umma_probe scripts/umma_smem_probe.cu && timeout 120 /tmp/umma_probe > gpurun_out/r02_probe.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest0.txt 2>&1
timeout 300 python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_base.json 2> gpurun_out/r02_c3_base.err
timeout 300 python bench.py --workload c3 --path 4 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_pair.json 2> gpurun_out/r02_c3_pair.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_pair -s 3 -c 1 -o gpurun_out/r02_k2pair -f python bench.py --workload c3 --path 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_ncu_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_gemm_topk -s 3 -c 1 -o gpurun_out/r02_k2base -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_ncu_base.log 2>&1
tail -3 gpurun_out/r02_pytest0.txt; cat gpurun_out/r02_probe.txt
