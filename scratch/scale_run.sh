#!/bin/bash
# 1/2/4/8-GPU weak-scaling run of bench.py on one box (driver-style launch)
cd "$(dirname "$0")/.."
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
done
for N in 1 2 4 8; do python - <<P
import json
try:
    d=json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
    print($N, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
except Exception as e:
    print($N, "failed", e)
P
done
