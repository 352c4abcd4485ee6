#!/bin/bash
# A/B of the x3 fused forward's issue order: selftest + tf32x3 tests on the default (interleaved) order, then the bench on both
mkdir -p gpurun_out
timeout 300 python scripts/selftest.py 2>&1 | grep -E "x3 fused|FAIL|rc "
timeout 900 python -m pytest tests/test_gpu_rows.py -k "tf32x3" -s -q --timeout 600 2>&1 | grep -E "B=128|passed|failed|rror"
for mode in interleaved sequential; do
  if [ $mode = sequential ]; then export PV_X3_SEQUENTIAL=1; else unset PV_X3_SEQUENTIAL; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scene-infer --no-side-tf32 --precision tf32x3 > gpurun_out/bench_ab_$mode.json 2> gpurun_out/bench_ab_$mode.err
  python - <<PY
import json
z=json.load(open("gpurun_out/bench_ab_$mode.json"))
print("$mode", "patches/s", round(z["value"],1), "ms/step", round(z["ms_per_step"],3), "resfront_fwd_x3", round(z["kernels"]["resfront_fwd_x3"]["ms_per_step"],3))
PY
done
