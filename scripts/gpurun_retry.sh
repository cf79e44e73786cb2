#!/bin/bash
# retries a gpurun call while the pod answers "busy" (exit code 3: nothing charged); usage: gpurun_retry.sh <timeout_s> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun ${GPURUN_GPUS:+--gpus $GPURUN_GPUS} --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
