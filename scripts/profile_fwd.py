"""Runs a few training steps of the p16t9c85r12 graph (for ncu captures): python scripts/profile_fwd.py [precision] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tempfile
import torch
import probav_b200 as pb
from probav_b200 import synth

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = pb.parseConfig(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cfg", "p16t9c85r12.cfg"))
m = pb.build_from_config(cfg, precision=prec)
L = pb.Losses((48, 48, 1))
d = tempfile.mkdtemp()
t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/c", d + "/l")
lr, hr, mask = synth.make_batch(128, seed=1, hr_zero_under_mask=True)
x, y, k = torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda(), torch.from_numpy(mask).cuda()
for _ in range(steps):
    t.trainStep(x, y, k, sync=False)
torch.cuda.synchronize()
print("done")
