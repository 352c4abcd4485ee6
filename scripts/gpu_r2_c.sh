#!/bin/bash
# round-2 visit C: smoke, whole GPU suite, both bench arms
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_r2c.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r2c.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_r2c.log 2>&1; tail -5 gpurun_out/pytest_r2c.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2c.err
python - <<PY
import json
z=json.load(open("gpurun_out/bench_r2c.json"))
print("patches/s", round(z["value"],1), "ms/step", round(z["ms_per_step"],3), "e2e", round(z["e2e"]["value"],1), z["clocks"], "launches", z["gpu_launches"])
print("side", z["single_pass_tf32"]); print("scene", z["scene_infer"]); print("cpu", z["cpu_baseline"])
r=z["roofline"]; print({k:r[k] for k in r if k!="note"})
for k,v in z["kernels"].items():
    if v["ms_per_step"]>0.02: print(f"{k:24s} {v['launches_per_step']:3d} {v['ms_per_step']:8.3f} ms  {v['tflops'] and round(v['tflops'],1)} exec {v['executed_tflops'] and round(v['executed_tflops'],1)}")
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2c.json 2>> gpurun_out/bench_r2c.err; cat gpurun_out/bench_ref_r2c.json | cut -c1-300
