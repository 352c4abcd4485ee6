"""Generates tests/golden/grad_b128_golden.npz: loss, cPSNR and per-tensor gradients of ONE full-size training batch
(cfg/p16t9c85r12, batch 128 = BASELINE configs[1]) from the fp64 CPU oracle (reference trainClass.py:126-131: forward,
shift-L1 loss, tape.gradient).  Run from the repo root: `python tests/golden/make_grad_b128_golden.py` (about 3 minutes on
8 cores; the batch is processed in chunks of 16 patches -- the loss is a mean over samples, so the per-chunk gradients
add up exactly).

Inputs are NOT stored: they are re-created by the seeded generators (`synth.make_batch(128, seed=SEED_DATA,
hr_zero_under_mask=False)`, i.e. raw HR under unclear pixels as the reference feeds it, and
`oracle.wdsr.init_params(specs, seed=SEED_W)`); a digest of each is stored so a drifting generator is detected.
PARITY UNPINNED w.r.t. TensorFlow (no TF in this image): this freezes the oracle, which is what the GPU engines are
compared with at the north_star's 1e-3 bar.
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.losses import OracleLosses  # noqa: E402
from oracle.step import loss_and_grads  # noqa: E402
from oracle.wdsr import OracleWDSR, init_params  # noqa: E402

NIR = (8075.2045, 3160.7272)
FULL = dict(scale=3, numFilters=32, kernelSize=(3, 3, 3), numResBlocks=12, expRate=8, decayRate=0.8,
            numImgLR=9, patchSizeLR=16, isGrayScale=True)
B, CHUNK, SEED_W, SEED_DATA = 128, 16, 100, 101


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pv_synth", os.path.join(ROOT, "proba-v_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    om = OracleWDSR(NIR[0], NIR[1], 6, **FULL)
    p = init_params(om.specs, seed=SEED_W, dtype=torch.float64)
    lr, hr, mask = synth.make_batch(B, seed=SEED_DATA, hr_zero_under_mask=False)
    ol = OracleLosses((48, 48, 1))
    t0 = time.time()
    loss = 0.0
    grads = {k: torch.zeros_like(v) for k, v in p.items()}
    cps, srs = [], []
    for s in range(0, B, CHUNK):
        sl = slice(s, s + CHUNK)
        l, g, sr, c = loss_and_grads(om, ol, p, torch.from_numpy(lr[sl]).double(), torch.from_numpy(hr[sl]).double(),
                                     torch.from_numpy(mask[sl]))
        w = CHUNK / B
        loss += float(l) * w
        for k in grads:
            grads[k] += g[k] * w
        cps.append(c.numpy())
        srs.append(sr.numpy())
        print(f"chunk {s // CHUNK + 1}/{B // CHUNK}: {time.time() - t0:.0f} s", flush=True)
    sr = np.concatenate(srs)
    _, idx, cnt, _ = ol.details("l1", torch.from_numpy(hr).double(), torch.from_numpy(mask), torch.from_numpy(sr))
    out = {"loss": np.float64(loss), "cpsnr": np.concatenate(cps), "best_shift": idx.numpy().astype(np.int32),
           "clear_count": cnt.numpy().astype(np.int32), "sr_head": sr[:4].astype(np.float32),
           "sr_absmax": np.float64(np.abs(sr).max()), "seed_w": SEED_W, "seed_data": SEED_DATA, "batch": B,
           "digest_inputs": digest(lr, hr, mask), "digest_weights": digest(*[p[k].numpy() for k in sorted(p)])}
    for k, v in grads.items():
        out["grad/" + k] = v.numpy().astype(np.float32)       # 535 267 floats in all
    np.savez_compressed(os.path.join(HERE, "grad_b128_golden.npz"), **out)
    print("loss", loss, "mean cPSNR", float(np.concatenate(cps).mean()), f"({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
