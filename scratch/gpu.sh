#!/bin/bash
# usage: scratch/gpu.sh <timeout-seconds> <log-name> '<command>'  -- retries while the pod answers busy (exit 3)
T=$1; NAME=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > gpurun_out/$NAME.call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (try $i)"; exit $rc; fi
  sleep 150
done
echo "gave up"; exit 3
