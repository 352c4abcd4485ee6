"""Throughput of ModelTrainer.fitTrainData (the reference-facing training loop: host arrays in, per-step host metric reads,
TensorBoard scalars) with and without the pinned-memory prefetch pipeline.  cfg/p16t9c85r12, batch 128, one GPU."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import probav_b200 as pb
from probav_b200 import synth

cfg = pb.parseConfig(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cfg", "p16t9c85r12.cfg"))
N, B = 4096, 128
X, y, msk = synth.make_batch(N, seed=7, hr_zero_under_mask=True)
out = {}
for prefetch in (False, True):
    m = pb.build_from_config(cfg, precision="tf32", seed=1)
    L = pb.Losses((48, 48, 1))
    d = tempfile.mkdtemp(prefix="pv_fit_")
    t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/ckpt", d + "/log", evalStep=10 ** 9)
    t.fitTrainData(X, [y, msk], B, 1, [X[:B], y[:B], msk[:B]], logEvery=0, maxSteps=4, prefetch=prefetch)      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t.fitTrainData(X, [y, msk], B, 1, [X[:B], y[:B], msk[:B]], logEvery=0, prefetch=prefetch)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["prefetch" if prefetch else "host_gather"] = {"steps": N // B, "patches_per_s": N / dt, "ms_per_step": dt / (N // B) * 1e3}
    t.close(); m.close()
print(json.dumps(out))
