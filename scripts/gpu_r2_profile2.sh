#!/bin/bash
# round-2 profile visit 2 (single-launch compensated conv): launch list of one steady-state tf32x3 train step, and an ncu --set full
# capture of two forward launches of the compensated 3x3x3 trunk conv (normConv_4 / _5 of the second step).
mkdir -p gpurun_out
L=$(python - <<PY
import os, sys, tempfile
sys.path.insert(0, os.getcwd())
import torch, probav_b200 as pb
from probav_b200 import synth, _lib
cfg = pb.parseConfig("cfg/p16t9c85r12.cfg")
m = pb.build_from_config(cfg, precision="tf32x3"); L = pb.Losses((48, 48, 1)); d = tempfile.mkdtemp()
t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/c", d + "/l")
lr, hr, mask = synth.make_batch(128, seed=1, hr_zero_under_mask=True)
x, y, k = torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda(), torch.from_numpy(mask).cuda()
t.trainStep(x, y, k, sync=False); torch.cuda.synchronize()
n0 = _lib.lib().pv_launch_count(); t.trainStep(x, y, k, sync=False); torch.cuda.synchronize()
print(_lib.lib().pv_launch_count() - n0)
PY
)
echo "tf32x3: $L launches per step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2 * L)) -c $L --csv --log-file gpurun_out/launches_r02_tf32x3_v3.csv \
    python scripts/profile_fwd.py tf32x3 3 > gpurun_out/ncu_ll_tf32x3_v3.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowconv3_tc -s 52 -c 2 -f -o gpurun_out/prof_r02_x3_conv3_v3 \
    python scripts/profile_fwd.py tf32x3 2 > gpurun_out/ncu_full_x3_v3.log 2>&1; echo "full conv3 rc=$?"
