#!/bin/bash
# round-2 profile visit: launch lists (ncu gpu__time_duration) of one steady-state train step of both tensor-core engines, and
# ncu --set full captures of their dominant kernels.  Usage: bash scripts/gpu_r2_profile.sh
mkdir -p gpurun_out
for PREC in tf32x3 tf32; do
  L=$(python - <<PY
import os, sys, tempfile
sys.path.insert(0, os.getcwd())
import torch, probav_b200 as pb
from probav_b200 import synth, _lib
cfg = pb.parseConfig("cfg/p16t9c85r12.cfg")
m = pb.build_from_config(cfg, precision="$PREC"); L = pb.Losses((48, 48, 1)); d = tempfile.mkdtemp()
t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/c", d + "/l")
lr, hr, mask = synth.make_batch(128, seed=1, hr_zero_under_mask=True)
x, y, k = torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda(), torch.from_numpy(mask).cuda()
t.trainStep(x, y, k, sync=False); torch.cuda.synchronize()
n0 = _lib.lib().pv_launch_count(); t.trainStep(x, y, k, sync=False); torch.cuda.synchronize()
print(_lib.lib().pv_launch_count() - n0)
PY
)
  echo "$PREC: $L launches per step"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2 * L)) -c $L --csv --log-file gpurun_out/launches_r02_${PREC}.csv \
      python scripts/profile_fwd.py $PREC 3 > gpurun_out/ncu_ll_${PREC}.log 2>&1; echo "launch list rc=$?"
done
# --set full of the dominant kernel classes of the tf32x3 step: the compensated conv's two passes (4 launches = blocks 5, 6 of the second step),
# then one launch each of the fused forward, the fused weight-gradient and the split-weight backward-data kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowconv3_tc -s 90 -c 4 -f -o gpurun_out/prof_r02_x3_conv3 \
    python scripts/profile_fwd.py tf32x3 2 > gpurun_out/ncu_full_x3a.log 2>&1; echo "full conv3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resfront" -s 40 -c 3 -f -o gpurun_out/prof_r02_x3_resfront \
    python scripts/profile_fwd.py tf32x3 2 > gpurun_out/ncu_full_x3b.log 2>&1; echo "full resfront rc=$?"
