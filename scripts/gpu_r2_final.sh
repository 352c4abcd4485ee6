#!/bin/bash
# last visit of round 2: smoke + whole GPU suite + both bench arms (gpu_r2_c.sh), then the launch list of one steady-state tf32x3 step
bash scripts/gpu_r2_c.sh
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
