#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> '<command>'   -- resubmits while the pod answers "busy" (exit 3), up to 40 min
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
