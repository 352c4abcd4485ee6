#!/bin/bash
# A/B of the data-parallel exchange at N GPUs: two overlapped buckets (default) vs one all-reduce after backward
N=${1:-8}
for V in 1 0; do
  PV_DP_OVERLAP=$V timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+V)) bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/scale_ab_$V.json 2> gpurun_out/scale_ab_$V.err
  python - <<PY
import json
z=json.load(open("gpurun_out/scale_ab_$V.json"))
print("N=$N overlap=$V", round(z["value"],1), "patches/s", round(z["ms_per_step"],4), "ms/step")
PY
done
