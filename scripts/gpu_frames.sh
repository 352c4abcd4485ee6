#!/bin/bash
# GPU visit for the T = 7 / 13 reducer tails on the row engines: parity tests, then a short T = 13 bench.
TAG=${1:-frames}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rows.py -q --timeout 300 -s -k "other_frame or t19 or small_graph or staged" > gpurun_out/pytest_${TAG}.log 2>&1
echo "pytest rc=$?"
grep -E "passed|failed|PvError|max rel err|worst grad|timed out|Error|assert" gpurun_out/pytest_${TAG}.log | sort | uniq -c | head -40
timeout 600 python bench.py --cfg cfg/p16t13c85r12.cfg --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_t13.json 2> gpurun_out/bench_${TAG}_t13.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG}_t13.err
python - <<PY
import json
z=json.load(open("gpurun_out/bench_${TAG}_t13.json"))
print("patches/s", round(z["value"],1), "ms/step", round(z["ms_per_step"],3), "e2e", round(z["e2e"]["value"],1), z["config"]["algorithmic_tflops"], z["clocks"])
for k,v in z["kernels"].items():
    if v["ms_per_step"]>0.02: print(f"{k:22s} {v['launches_per_step']:3d} {v['ms_per_step']:8.3f} ms  {v['tflops'] and round(v['tflops'],1)}")
PY
