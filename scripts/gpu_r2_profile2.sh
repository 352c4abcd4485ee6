#!/bin/bash
# round-2 profile visit 2 (final kernels): launch list of one steady-state tf32x3 train step, and ncu --set full captures of the
# compensated 3x3x3 trunk conv (forward and data gradient) and of the fused expand/decay kernels, all from the second step.
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
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2 * L)) -c $L --csv --log-file gpurun_out/launches_r02_tf32x3_final.csv \
    python scripts/profile_fwd.py tf32x3 3 > gpurun_out/ncu_ll_tf32x3_final.log 2>&1; echo "launch list rc=$?"
# per step the conv3 kernel runs 16 forward launches (MODE 1) and 16 data gradients (MODE 2); the fused kernels 12 forward + 12 x (weight, data)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowconv3_tc -s 36 -c 2 -f -o gpurun_out/prof_r02_x3_conv3_fwd \
    python scripts/profile_fwd.py tf32x3 2 > gpurun_out/ncu_full_x3_a.log 2>&1; echo "full conv3 fwd rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowconv3_tc -s 56 -c 2 -f -o gpurun_out/prof_r02_x3_conv3_dgrad \
    python scripts/profile_fwd.py tf32x3 2 > gpurun_out/ncu_full_x3_b.log 2>&1; echo "full conv3 dgrad rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resfront -s 40 -c 2 -f -o gpurun_out/prof_r02_x3_resfront_fwd \
    python scripts/profile_fwd.py tf32x3 2 > gpurun_out/ncu_full_x3_c.log 2>&1; echo "full resfront fwd rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:resfront -s 54 -c 2 -f -o gpurun_out/prof_r02_x3_resfront_bwd \
    python scripts/profile_fwd.py tf32x3 2 > gpurun_out/ncu_full_x3_d.log 2>&1; echo "full resfront bwd rc=$?"
