#!/bin/bash
# round-2 GPU call 10: cooperative-launch check for the one-CTA K2, small-collection breakdowns, K1 trace on tiny shards, select stage
set -x
cd "$GRAFT_REPO_ROOT"
timeout 120 python scripts/small_batch.py 10000 16 > gpurun_out/r02_small16_plain.txt 2>&1
cat gpurun_out/r02_small16_plain.txt
YRB_K2_FUSE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small16u_launches.csv python scripts/small_batch.py 10000 16 > gpurun_out/r02_ncu_small16u.log 2>&1
YRB_K1_TRACE=1 timeout 120 python scripts/k1_small.py 1000 > gpurun_out/r02_k1trace_1000.txt 2>&1
YRB_K1_TRACE=1 timeout 120 python scripts/k1_small.py 125000 > gpurun_out/r02_k1trace_125k.txt 2>&1
timeout 120 python scripts/k1_small.py 1000 > gpurun_out/r02_k1small_1000.txt 2>&1
timeout 120 python scripts/k1_small.py 125000 > gpurun_out/r02_k1small_125k.txt 2>&1
timeout 300 python bench.py --workload t10mbs --steps 50 --warmup 5 --no-cpu --no-parity > gpurun_out/r02_t10mbs2.json 2> gpurun_out/r02_t10mbs2.err
timeout 300 python bench.py --workload c5s --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_c5s3.json 2> gpurun_out/r02_c5s3.err
timeout 300 python scripts/crossover.py 1000 10000 > gpurun_out/r02_crossover2.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest5.txt 2>&1
tail -4 gpurun_out/r02_pytest5.txt
