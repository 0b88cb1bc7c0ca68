#!/bin/bash
# round-2 GPU call 6: the new default bench at N=1 (timed), smoke, sharded tests
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_sh2.txt 2>&1
tail -3 gpurun_out/r02_pytest_sh2.txt
( time timeout 900 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_smoke.txt 2>&1
tail -5 gpurun_out/r02_smoke.txt
( time timeout 1200 python bench.py --steps 100 --warmup 10 ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -5 gpurun_out/r02_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
nproc; free -g | head -2
