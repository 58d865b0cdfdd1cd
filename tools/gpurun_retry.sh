#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3, nothing charged).
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] '<command>'
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout "$T" $G -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
