#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> [--gpus N] <command...>   -- retries gpurun while the pod answers "busy" (exit 3, not charged)
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift; shift; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" $G -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
