#!/bin/bash
# round-2 GPU call 3: branch-free epilogue (one-CTA and pair), launch list of the pair path
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_store.py -m gpu -x -q > gpurun_out/r02_pytest2.txt 2>&1
timeout 300 python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_base3.json 2> gpurun_out/r02_c3_base3.err
timeout 300 python bench.py --workload c3 --path 4 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_pair3.json 2> gpurun_out/r02_c3_pair3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_pair3_launches.csv python bench.py --workload c3 --path 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_ncu_pair3l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_gemm_topk_pair -s 3 -c 1 -o gpurun_out/r02_k2pair3 -f python bench.py --workload c3 --path 4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_ncu_pair3.log 2>&1
tail -3 gpurun_out/r02_pytest2.txt
