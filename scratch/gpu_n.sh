#!/bin/bash
# usage: scratch/gpu_n.sh <gpus> <timeout-seconds> <log-name> '<command>'  -- retries while the pod answers busy (exit 3)
G=$1; T=$2; NAME=$3; shift 3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > gpurun_out/$NAME.call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (try $i)"; exit $rc; fi
  sleep 180
done
echo "gave up"; exit 3
