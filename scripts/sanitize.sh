#!/bin/bash
# compute-sanitizer over the tensor-core kernels: the device self-test (every tcgen05 kernel configuration, B = 3) and one train step
# of the 2-block graph on both tensor-core engines.  Usage: bash scripts/sanitize.sh [tag]   (logs under gpurun_out/)
TAG=${1:-r02}
mkdir -p gpurun_out
cat > /tmp/pv_san_step.py <<'PY'
import os, sys, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import probav_b200 as pb
from probav_b200 import synth
cfg = dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=2, expRate=8, decayRate=0.8, numImgLR=9, patchSizeLR=16, isGrayScale=True)
for prec in ("tf32", "tf32x3"):
    m = pb.WDSRConv3D("superResolutionNet", "NIR", 8075.2045, 3160.7272, 6).build(**cfg, precision=prec)
    L = pb.Losses((48, 48, 1)); d = tempfile.mkdtemp()
    t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/c", d + "/l")
    lr, hr, mask = synth.make_batch(2, seed=1)
    print(prec, t.trainStep(lr, hr, mask))
PY
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/selftest.py > gpurun_out/sanitizer_${tool}_selftest_${TAG}.log 2>&1
  echo "$tool selftest rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc " gpurun_out/sanitizer_${tool}_selftest_${TAG}.log | tail -3
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/pv_san_step.py > gpurun_out/sanitizer_${tool}_step_${TAG}.log 2>&1
  echo "$tool step rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^tf32" gpurun_out/sanitizer_${tool}_step_${TAG}.log | tail -4
done
