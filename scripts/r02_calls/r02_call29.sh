#!/bin/bash
# final check of the default bench line after the side-leg guard
set -x
cd /root/repo
timeout 200 python bench.py --steps 40 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
echo rc=$?
