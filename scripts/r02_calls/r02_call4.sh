#!/bin/bash
# round-2 GPU call 4: fused in-kernel sampling, pair default, MAX_TOPS=4; crossover table; C5 shard
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_store.py tests/test_gpu_search.py -m gpu -x -q > gpurun_out/r02_pytest3.txt 2>&1
tail -3 gpurun_out/r02_pytest3.txt
timeout 300 python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_fused.json 2> gpurun_out/r02_c3_fused.err
YRB_K2_FUSE=0 timeout 300 python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_unfused.json 2> gpurun_out/r02_c3_unfused.err
YRB_K2_PAIR=0 timeout 300 python bench.py --workload c3 --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3_onecta.json 2> gpurun_out/r02_c3_onecta.err
timeout 300 python bench.py --workload c3s --steps 100 --warmup 10 --no-cpu > gpurun_out/r02_c3s.json 2> gpurun_out/r02_c3s.err
timeout 600 python bench.py --workload c5s --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_c5s.json 2> gpurun_out/r02_c5s.err
timeout 600 python scripts/crossover.py > gpurun_out/r02_crossover.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c3_fused_launches.csv python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_ncu_c3f.log 2>&1
cat gpurun_out/r02_crossover.txt
