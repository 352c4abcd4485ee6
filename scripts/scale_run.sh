#!/bin/bash
# weak-scaling run on one 8-GPU box: N = 1, 2, 4, 8 back to back (bench.py contract)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  python - <<PY
import json
try:
    z=json.load(open("gpurun_out/scale_n$N.json"))
    print("N=$N", round(z["value"],1), "patches/s", round(z["ms_per_step"],3), "ms/step  e2e", round(z["e2e"]["value"],1), z["clocks"])
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/scale_n$N.err").read()[-800:])
PY
done
