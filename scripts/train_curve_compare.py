"""Do the engines TRAIN alike?  K Nadam steps of cfg/p16t9c85r12 (batch 128, fresh synthetic batch per step, same seeds for every engine)
on the error-compensated tensor-core engine, the single-pass one and the fp32 CUDA-core engine; prints the loss / cPSNR trajectories side by
side and the relative distance of the final weights.   python scripts/train_curve_compare.py [steps] [engines, comma separated]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import probav_b200 as pb
from probav_b200 import synth

K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
engines = (sys.argv[2] if len(sys.argv) > 2 else "fp32_rows,tf32x3,tf32").split(",")
FULL = dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=12, expRate=8, decayRate=0.8, numImgLR=9, patchSizeLR=16, isGrayScale=True)
NIR = (8075.2045, 3160.7272)
curves, weights = {}, {}
for prec in engines:
    m = pb.WDSRConv3D("n", "NIR", NIR[0], NIR[1], 6).build(**FULL, precision=prec, seed=7)
    L = pb.Losses((48, 48, 1)); d = tempfile.mkdtemp()
    t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/c", d + "/l")
    c = []
    for k in range(K):
        x, y, msk = synth.make_batch(128, seed=50_000 + k, hr_zero_under_mask=False)
        c.append(t.trainStep(x, y, msk))
    curves[prec] = np.array(c)
    weights[prec] = np.concatenate([v.ravel() for v in m.get_weights().values()])
    t.close(); m.close()
ref = engines[0]
print(f"{'step':>5s} " + " ".join(f"{e + ' loss':>16s} {e + ' dB':>12s}" for e in engines))
for k in sorted(set(list(range(0, K, max(1, K // 10))) + [K - 1])):
    print(f"{k:5d} " + " ".join(f"{curves[e][k, 0]:16.4f} {curves[e][k, 1]:12.4f}" for e in engines))
for e in engines[1:]:
    dl = np.abs(curves[e][:, 0] - curves[ref][:, 0]) / curves[ref][:, 0]
    dw = np.linalg.norm(weights[e] - weights[ref]) / np.linalg.norm(weights[ref])
    print(f"{e} vs {ref}: loss rel. difference first step {dl[0]:.2e}, max over steps 0-9 {dl[:10].max():.2e}, max over all {K} steps {dl.max():.2e}, "
          f"mean over the last 20 {dl[-20:].mean():.2e}; final weights |dw| / |w| = {dw:.2e}")
