#!/bin/bash
# round-2 GPU call 24 (1 GPU): refresh the committed ncu evidence on the final code
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_plain_default.json 2> gpurun_out/r02_plain_default.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_default_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_default.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_scan_topk -s 6 -c 1 -o gpurun_out/r02_k1_t10m -f python bench.py --workload t10m --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_gemm_topk_pair -s 4 -c 1 -o gpurun_out/r02_k2pair_final -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_k2f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_segments -s 4 -c 1 -o gpurun_out/r02_k3_c3 -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_k3.log 2>&1
