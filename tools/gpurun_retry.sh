#!/bin/bash
# gpurun with retries on "busy" (exit 3): tools/gpurun_retry.sh <timeout> '<command>'  [extra gpurun args, e.g. --gpus 2]
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" -- "$CMD"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
