#!/bin/bash
# round-2 visit B: the error-compensated engine (precision tf32x3): self-test, parity, full-batch golden gradients, bench
mkdir -p gpurun_out
timeout 300 python scripts/selftest.py > gpurun_out/selftest_r2b.log 2>&1; tail -30 gpurun_out/selftest_r2b.log
timeout 900 python -m pytest tests/test_gpu_rows.py -k "tf32x3" -s -q --timeout 600 > gpurun_out/pytest_r2b.log 2>&1
grep -E "B=128|passed|failed|rror|max rel err|worst grad|assert" gpurun_out/pytest_r2b.log | head -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision tf32x3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2b.err
python - <<PY
import json
z=json.load(open("gpurun_out/bench_r2b.json"))
print("patches/s", round(z["value"],1), "ms/step", round(z["ms_per_step"],3), "e2e", round(z["e2e"]["value"],1), z["clocks"], z["scene_infer"] and round(z["scene_infer"]["value"],1))
for k,v in z["kernels"].items():
    if v["ms_per_step"]>0.02: print(f"{k:24s} {v['launches_per_step']:3d} {v['ms_per_step']:8.3f} ms  {v['tflops'] and round(v['tflops'],1)}")
PY
