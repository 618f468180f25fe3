#!/bin/bash
# usage: [GPURUN_ARGS='--gpus 2'] scripts/gpurun_retry.sh <timeout_s> '<command>'   -- resubmits while the pod answers "busy" (exit 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPURUN_ARGS:-} --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
