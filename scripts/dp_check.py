"""2-GPU check of the data-parallel fit loop (run under torchrun): train.py's path with the prefetch pipeline, staged
backward + bucketed all-reduce; every rank must end with bit-identical weights, rank 0 alone writes the checkpoint and the
event file, and the result must agree with a single-GPU run over the same global batches (to fp32 reduction-order noise)."""
import glob
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import probav_b200 as pb
from probav_b200 import parallel, synth, tfckpt

rank, ws, local = parallel.init_from_env()
torch.cuda.set_device(local)
cfg = dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=2, expRate=8, decayRate=0.8, numImgLR=9, patchSizeLR=16,
           isGrayScale=True)
X, y, msk = synth.make_batch(96, seed=5, hr_zero_under_mask=True)
d = tempfile.mkdtemp(prefix=f"pv_dp_{rank}_") if rank else os.environ.get("PV_DP_DIR", tempfile.mkdtemp(prefix="pv_dp_0_"))


def run(world_mode: bool):
    m = pb.WDSRConv3D("n", "NIR", 8075.2045, 3160.7272, 6).build(**cfg, seed=3, precision=os.environ.get("PV_DP_PRECISION", "tf32x3"), device=local)
    L = pb.Losses((48, 48, 1))
    sub = "dp" if world_mode else "single"
    t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), f"{d}/{sub}/ckpt", f"{d}/{sub}/log", evalStep=2)
    t.fitTrainData(X, [y, msk], 32, 1, [X[:32], y[:32], msk[:32]], valSteps=1, saveBestOnly=False, logEvery=0)
    w = m.param_arena().clone()
    t.close()
    return w, t


w_dp, t = run(True)
ws_list = [torch.empty_like(w_dp) for _ in range(ws)]
dist.all_gather(ws_list, w_dp)
same = all(torch.equal(ws_list[0], o) for o in ws_list)
files = sorted(os.path.basename(f) for f in glob.glob(f"{d}/dp/ckpt/*"))
if rank == 0:
    assert same, "ranks diverged"
    assert t.step == 3 and "checkpoint" in files and any(f.endswith(".index") for f in files), files
    assert glob.glob(f"{d}/dp/log/events.out.tfevents.*")
    print("[dp_check] ranks identical after 3 global steps; rank 0 wrote", files)
else:
    assert not files and not glob.glob(f"{d}/dp/log/events.out.tfevents.*"), "only rank 0 may write"
dist.barrier()
# single-GPU reference over the same global batches: temporarily pretend world size 1 (rank 0 only)
if rank == 0:
    real_world = parallel.world
    parallel.world = lambda: (0, 1)
    try:
        w_1, _ = run(False)
    finally:
        parallel.world = real_world
    err = float((w_1 - w_dp).abs().max()), float(w_1.abs().max())
    print(f"[dp_check] max |w_dp - w_single| = {err[0]:.3e} (max |w| {err[1]:.3e})")
    assert err[0] < 2e-4 * err[1]
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("[dp_check] OK")
