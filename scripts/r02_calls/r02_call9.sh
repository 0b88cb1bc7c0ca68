#!/bin/bash
# round-2 GPU call 9: per-kernel breakdowns (one t10mb shard; 10k x 16-query batch; single query on 1000 rows)
set -x
cd "$GRAFT_REPO_ROOT"
timeout 300 python bench.py --workload t10mbs --steps 50 --warmup 5 --no-cpu --no-parity > gpurun_out/r02_t10mbs.json 2> gpurun_out/r02_t10mbs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_t10mbs_launches.csv python bench.py --workload t10mbs --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_t10mbs.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small16_launches.csv python scripts/small_batch.py 10000 16 > gpurun_out/r02_ncu_small16.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small1_launches.csv python scripts/small_batch.py 1000 1 > gpurun_out/r02_ncu_small1.log 2>&1
timeout 300 python bench.py --workload c5s --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_c5s2.json 2> gpurun_out/r02_c5s2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c5s_launches.csv python bench.py --workload c5s --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_c5s.log 2>&1
