#!/bin/bash
# usage: gpu_retry.sh <timeout-seconds> <command string> [extra gpurun flags]
# retries gpurun while the pod answers "busy" (exit 3), at most 40 minutes
T=$1; CMD=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
