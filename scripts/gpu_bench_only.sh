#!/bin/bash
# GPU visit: gradient parity of the row engines + a short bench with the per-kernel breakdown.  Usage: bash scripts/gpu_bench_only.sh tag
TAG=${1:-b}
timeout 600 python -m pytest tests/test_gpu_rows.py -q --timeout 300 -x -k "gradients or staged" 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scene-infer > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
z=json.load(open("gpurun_out/bench_${TAG}.json"))
print("patches/s", round(z["value"],1), "ms/step", round(z["ms_per_step"],3), "e2e", round(z["e2e"]["value"],1), z["clocks"])
for k,v in z["kernels"].items():
    if v["ms_per_step"]>0.02: print(f"{k:22s} {v['launches_per_step']:3d} {v['ms_per_step']:8.3f} ms  {v['tflops'] and round(v['tflops'],1)}")
PY
