#!/bin/bash
# round-2 GPU call 33 (1 GPU): smoke() on the final tree
set -x
cd "$GRAFT_REPO_ROOT"
( time timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r02_smoke_final.txt 2>&1
echo rc=$?
tail -5 gpurun_out/r02_smoke_final.txt
