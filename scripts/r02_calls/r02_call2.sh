#!/bin/bash
# round-2 GPU call 2: K2 with elect_one issue loops (one-CTA and pair), ncu of both
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_store.py -m gpu -x -q > gpurun_out/r02_pytest1.txt 2>&1
timeout 300 python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_base2.json 2> gpurun_out/r02_c3_base2.err
timeout 300 python bench.py --workload c3 --path 4 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_pair2.json 2> gpurun_out/r02_c3_pair2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_gemm_topk_pair -s 3 -c 1 -o gpurun_out/r02_k2pair2 -f python bench.py --workload c3 --path 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_ncu_pair2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_gemm_topk -s 3 -c 1 -o gpurun_out/r02_k2base2 -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_ncu_base2.log 2>&1
tail -3 gpurun_out/r02_pytest1.txt
