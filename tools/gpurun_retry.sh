#!/bin/bash
# tools/gpurun_retry.sh <timeout-seconds> <command...>: gpurun, retried while the pod answers "busy" (rc 3)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
