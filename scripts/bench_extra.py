"""Extra measurements quoted in DESIGN.md (not part of the bench.py contract):
   (1) shift-loss kernel at a large batch (HBM roofline view);  (2) 384x384 scene inference (BASELINE configs[3])."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import probav_b200 as pb
from probav_b200 import synth

out = {}
# ---- (1) shift loss, B = 65536 samples of 48x48 (1.36 GB of HR + SR + mask)
B = 65536
g = torch.Generator(device="cuda").manual_seed(0)
hr = torch.round(torch.rand(B, 48, 48, 1, device="cuda", generator=g) * 4000 + 6000)
sr = hr.roll((1, -2), (1, 2)) + torch.randn(B, 48, 48, 1, device="cuda", generator=g) * 40
mask = torch.rand(B, 48, 48, 1, device="cuda", generator=g) > 0.08
L = pb.Losses((48, 48, 1))
for kind, grad in (("l1", False), ("l1", True), ("sobel_l1_mix", True)):
    for _ in range(3):
        L.evaluate(kind, hr, mask, sr, want_grad=grad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        L.evaluate(kind, hr, mask, sr, want_grad=grad)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    by = B * 2304 * (9 + (4 if grad else 0))
    out[f"shift_loss_{kind}{'_bwd' if grad else ''}_B65536"] = {"ms": ms, "GBps": by / ms / 1e6, "frac_of_hbm_6542.7": by / ms / 1e6 / 6542.7,
                                                             "samples_per_s": B / ms * 1e3}
del hr, sr, mask
# ---- (2) scene inference: 9 x 128 x 128 LR -> 384 x 384, 64 patches per scene, clip + round + stitch on device
cfg = pb.parseConfig(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cfg", "p16t9c85r12.cfg"))
for prec in ("tf32", "fp32"):
    m = pb.build_from_config(cfg, precision=prec)
    ns = 32
    lr, hrs, msk = synth.make_scene(ns, seed=3)
    m.predict_from_scenes(lr)                 # warm-up at the full batch: the activation pools are sized on first use
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        srs = m.predict_from_scenes(lr)          # host in, host out (H2D + patching + forward + resolve + stitch + D2H)
    dt = (time.perf_counter() - t0) / reps
    c = pb.Losses((384, 384, 1)).shiftCompensatedcPSNR(hrs, msk, srs)
    out[f"scene_infer_{prec}"] = {"scenes_per_s_e2e_host": ns / dt, "ms_per_scene": dt / ns * 1e3, "mean_cpsnr_random_weights": float(np.mean(c))}
    m.close()
print(json.dumps(out, indent=1))
