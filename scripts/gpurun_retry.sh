#!/bin/bash
# usage: scripts/gpurun_retry.sh OUTFILE [gpurun args...] — retries while the pod answers "transient" (nothing charged)
out="$1"; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
