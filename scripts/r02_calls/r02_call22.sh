#!/bin/bash
# round-2 GPU call 22 (1 GPU): pair kernel with software-pipelined TMEM loads — parity + C3 / c5s timing
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_config_shapes.py -m gpu -x -q > gpurun_out/r02_pytest_pipe.txt 2>&1
tail -4 gpurun_out/r02_pytest_pipe.txt
for i in 1 2; do timeout 300 python bench.py --workload c3 --steps 200 --warmup 20 --no-cpu --no-parity > gpurun_out/r02_c3_pipe$i.json 2> gpurun_out/r02_c3_pipe$i.err; done
timeout 300 python bench.py --workload t10mbs --steps 100 --warmup 10 --no-cpu --no-parity > gpurun_out/r02_t10mbs_pipe.json 2> gpurun_out/r02_t10mbs_pipe.err
