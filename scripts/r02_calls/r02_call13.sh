#!/bin/bash
# round-2 GPU call 13: threshold tests, default bench line, ncu launch list of the same command, ncu --set full of K1 (t10m) and K2 pair (c3)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_where.py tests/test_gpu_store.py -m gpu -q -x > gpurun_out/r02_pytest8.txt 2>&1
tail -6 gpurun_out/r02_pytest8.txt
timeout 1200 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_default2.json 2> gpurun_out/r02_bench_default2.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_default_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_default.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_scan_topk -s 6 -c 1 -o gpurun_out/r02_k1_t10m -f python bench.py --workload t10m --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_gemm_topk_pair -s 4 -c 1 -o gpurun_out/r02_k2pair_final -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_k2f.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:ingest_stream -s 3 -c 1 -o gpurun_out/r02_k5 -f python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_k5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
