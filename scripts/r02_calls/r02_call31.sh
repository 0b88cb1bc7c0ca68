#!/bin/bash
# round-2 GPU call 31 (1 GPU): launch list of the final default command (t10m + also{c4,t10mb,t10mq,c3})
set -x
cd "$GRAFT_REPO_ROOT"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_default_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > gpurun_out/r02_ncu_default.log 2>&1
echo rc=$?
