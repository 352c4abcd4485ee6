"""CPU study (no GPU): gradient error of the tensor-core engine's NUMERICS MODEL (oracle/tf32_model.py) against the fp64
oracle, for a grid of operand-quantisation modes.  Usage:  python scripts/tf32_study.py [B] [mode-spec ...]
  mode-spec = comma-separated key=value over act / wt / grad / stream / gstream, e.g.  act=rn,wt=x2,grad=x2
Prints, per mode, the per-tensor max-norm relative gradient error (the metric of tests/test_gpu_rows.py) -- worst, and the
worst five tensors -- plus loss and SR errors.  Results are summarised in profiles/r02_tf32_numerics_study.md.
"""
import importlib.util
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.losses import OracleLosses  # noqa: E402
from oracle.step import loss_and_grads  # noqa: E402
from oracle.tf32_model import TensorCoreModel  # noqa: E402
from oracle.wdsr import OracleWDSR, init_params  # noqa: E402

NIR = (8075.2045, 3160.7272)
FULL = dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=12, expRate=8, decayRate=0.8,
            numImgLR=9, patchSizeLR=16, isGrayScale=True)
CHUNK = 16


def run(model, p, lr, hr, mask):
    B = lr.shape[0]
    ol = OracleLosses((48, 48, 1))
    loss, grads, srs = 0.0, None, []
    for s in range(0, B, CHUNK):
        sl = slice(s, min(B, s + CHUNK))
        n = sl.stop - sl.start
        l, g, sr, _ = loss_and_grads(model, ol, p, torch.from_numpy(lr[sl]).double(), torch.from_numpy(hr[sl]).double(), torch.from_numpy(mask[sl]))
        loss += float(l) * n / B
        if grads is None:
            grads = {k: v * (n / B) for k, v in g.items()}
        else:
            for k in grads:
                grads[k] += g[k] * (n / B)
        srs.append(sr.numpy())
    return loss, {k: v.numpy() for k, v in grads.items()}, np.concatenate(srs)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    specs = sys.argv[2:] or ["act=rn,wt=rn,grad=rn,stream=rn,gstream=rn"]
    spec = importlib.util.spec_from_file_location("pv_synth", os.path.join(ROOT, "proba-v_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    om = OracleWDSR(NIR[0], NIR[1], 6, **FULL)
    p = init_params(om.specs, seed=100, dtype=torch.float64)
    lr, hr, mask = synth.make_batch(128, seed=101, hr_zero_under_mask=False)
    lr, hr, mask = lr[:B], hr[:B], mask[:B]
    gold = os.path.join(ROOT, "tests", "golden", "grad_b128_golden.npz")
    if B == 128 and os.path.exists(gold):
        z = np.load(gold)
        ref_loss, ref_g, ref_sr = float(z["loss"]), {k[5:]: z[k] for k in z.files if k.startswith("grad/")}, None
    else:
        ref_loss, ref_g, ref_sr = run(om, p, lr, hr, mask)
    for s in specs:
        mode = dict(kv.split("=") for kv in s.split(",") if kv)
        tm = TensorCoreModel(NIR[0], NIR[1], 6, **FULL, mode=mode)
        t0 = time.time()
        loss, g, sr = run(tm, p, lr, hr, mask)
        errs = []
        for k, r in ref_g.items():
            den = np.abs(r).max()
            if den == 0:
                continue
            errs.append((float(np.abs(g[k] - r).max() / den), float(np.linalg.norm(g[k] - r) / (np.linalg.norm(r) + 1e-300)), k))
        errs.sort(reverse=True)
        sr_err = float(np.abs(sr - ref_sr).max() / np.abs(ref_sr).max()) if ref_sr is not None else float("nan")
        print(f"B={B} mode {tm.mode}: loss rel {abs(loss - ref_loss) / ref_loss:.2e}, SR {sr_err:.2e}, worst grad max-norm {errs[0][0]:.2e} "
              f"(median {np.median([e[0] for e in errs]):.2e}; worst L2-rel {max(e[1] for e in errs):.2e}); n>1e-3: {sum(e[0] > 1e-3 for e in errs)}/{len(errs)}  [{time.time() - t0:.0f} s]")
        print("    " + "; ".join(f"{k} {a:.1e}" for a, _, k in errs[:5]), flush=True)


if __name__ == "__main__":
    main()
