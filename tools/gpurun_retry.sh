#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   — retries while the pod answers "busy" (exit code 3, nothing charged)
T=$1; shift
for i in $(seq 1 200); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 15
done
exit 3
