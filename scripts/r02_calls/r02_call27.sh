#!/bin/bash
# round-2 GPU call 27 (1 GPU): the final code — whole suite, smoke, default bench line
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_pytest_final5.txt 2>&1
tail -4 gpurun_out/r02_pytest_final5.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final3.txt 2>&1; tail -1 gpurun_out/r02_smoke_final3.txt
timeout 900 python bench.py --gpus 1 --steps 100 --warmup 10 > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err
tail -2 gpurun_out/r02_bench_final2.err
