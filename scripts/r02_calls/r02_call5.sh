#!/bin/bash
# round-2 GPU call 5: sharded index (several shards on one GPU) + the whole GPU suite
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_pytest_sh.txt 2>&1
tail -30 gpurun_out/r02_pytest_sh.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.txt 2>&1
tail -5 gpurun_out/r02_pytest4.txt
