#!/bin/bash
# tools/gpurun_retry.sh <out-file> <gpurun args...>: retries while the pod answers "busy" (rc 3), at most ~40 min
out=$1; shift
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
