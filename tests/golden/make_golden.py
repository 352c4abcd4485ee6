"""Regenerates the golden vectors under tests/golden/ from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).

The reference (TensorFlow 2.x + tensorflow-addons) cannot be imported in this image, and it ships no golden vectors
of its own, so these fixtures freeze the *oracle's* outputs on seeded synthetic inputs: they pin the oracle against
silent drift and give the GPU parity tests a file-based target.  PARITY UNPINNED w.r.t. TensorFlow itself.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.losses import OracleLosses  # noqa: E402
from oracle.wdsr import OracleWDSR, init_params  # noqa: E402

NIR = (8075.2045, 3160.7272)
SMALL = dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=2, expRate=8, decayRate=0.8,
             numImgLR=9, patchSizeLR=16, isGrayScale=True)


def shift_loss_golden():
    g = torch.Generator().manual_seed(1234)
    B = 6
    hr = torch.round(torch.rand(B, 48, 48, 1, generator=g, dtype=torch.float64) * 4000 + 6000)
    sr = (hr.roll((2, -1), (1, 2)) + torch.randn(B, 48, 48, 1, generator=g, dtype=torch.float64) * 45).float().double()
    mask = torch.rand(B, 48, 48, 1, generator=g) > 0.12
    mask[0] = True                      # one all-clear sample
    hr[1] = hr[1] * mask[1]             # one sample with HR zeroed under the mask
    L = OracleLosses((48, 48, 1))
    out = {"hr": hr.numpy().astype(np.float32), "sr": sr.numpy().astype(np.float32), "mask": mask.numpy()}
    for kind in ("l1", "l2"):
        best, idx, cnt, stack = L.details(kind, hr, mask, sr)
        out[f"loss_{kind}"] = best.numpy()
        out[f"best_shift_{kind}"] = idx.numpy().astype(np.int32)
        out[f"clear_count_{kind}"] = cnt.numpy().astype(np.int32)
        out[f"stack_{kind}"] = stack.T.numpy()
        srg = sr.clone().requires_grad_(True)
        (L.shiftCompensatedL1Loss if kind == "l1" else L.shiftCompensatedL2Loss)(hr, mask, srg).backward()
        out[f"dsr_{kind}"] = srg.grad.numpy()
    best, idx, _, stack = L.details("l1edge", hr, mask, sr)
    out["loss_l1edge"] = best.numpy()
    out["best_shift_l1edge"] = idx.numpy().astype(np.int32)
    out["cpsnr"] = L.shiftCompensatedcPSNR(hr, mask, sr).numpy()
    np.savez_compressed(os.path.join(HERE, "shift_loss_golden.npz"), **out)


def wdsr_small_golden():
    from probav_b200_synth import make_batch
    om = OracleWDSR(NIR[0], NIR[1], 6, **SMALL)
    seed = 7
    p = init_params(om.specs, seed=seed)
    lr, _, _ = make_batch(2, seed=8)
    with torch.no_grad():
        sr = om.forward(p, torch.from_numpy(lr).double()).numpy()
    np.savez_compressed(os.path.join(HERE, "wdsr_small_golden.npz"), lr=lr, sr=sr, weight_seed=seed)


if __name__ == "__main__":
    # synth.py is pure numpy; import it without importing the package (which needs no GPU either, but keep this light)
    import importlib.util
    spec = importlib.util.spec_from_file_location("probav_b200_synth", os.path.join(HERE, "..", "..", "proba-v_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["probav_b200_synth"] = mod
    shift_loss_golden()
    wdsr_small_golden()
    print("golden vectors written to", HERE)
