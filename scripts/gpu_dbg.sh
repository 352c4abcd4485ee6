#!/bin/bash
# timing experiments: bench kernel table under debug knobs.  Usage: bash scripts/gpu_dbg.sh VAR "v1 v2 ..." kernel-regex
VAR=$1; VALS=$2; RE=$3
for v in $VALS; do
  env $VAR=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/dbg.json 2>/dev/null
  python - <<PY
import json,re
z=json.load(open("gpurun_out/dbg.json"))
print("$VAR=$v", round(z["ms_per_step"],3), {k: round(v["ms_per_step"],3) for k,v in z["kernels"].items() if re.search("$RE",k)})
PY
done
