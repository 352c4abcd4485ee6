"""CPU study (no GPU): does tcgen05's TRUNCATING fp32 accumulation explain what is left of the tf32x3 engine's gradient error?

oracle/tf32_model.py models the operand formats with exact accumulation and predicts 1.5e-4 for the final engine (hardware: 4.3e-4,
profiles/r02_tf32_numerics_study.md).  This script replaces the FORWARD values of the model's tensor-core layers by an emulation of the
kernels' MMA chains: every MMA adds an exactly computed partial product to an fp32 accumulator that is then rounded TOWARD ZERO; the
chains, their lengths and what is summed on the CUDA cores afterwards (round to nearest) follow the kernels:

  3x3x3 conv (conv3_tc.cu MODE 1)   three accumulators (one per dw tap), each 9 (dt, dh) x 2 K = 16 steps over the fp16 hi halves;
                                    the correction accumulator is 2^-11 of the size (its truncations are ignored here);
                                    out = rn(rn(rn(Q1 + Q0) + Q2) + C) + bias
  fused block (resblock_x3_tc.cu)   E_q: 12 tf32 K = 8 steps (x_hi w_hi, x_lo w_hi, x_hi w_lo) + 1 bias step; fp16 pair of relu(E);
                                    D: 16 K = 16 steps over E16 Wd_hi, corrections exact, + bias

The backward pass is the model's (gradients flow through the exact-accumulation graph; only the forward VALUES, hence the L1 signs and
ReLU masks, change).   python scripts/accum_trunc_study.py [B] [trunc|rn]"""
import importlib.util
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.tf32_model import TensorCoreModel, q_rn, q_tr  # noqa: E402
from oracle.wdsr import OracleWDSR, init_params, wn_kernel  # noqa: E402
from scripts.tf32_study import FULL, NIR, run  # noqa: E402

MODE = dict(act="f16p", wt="f16p", grad="rn", stream="none", gstream="rn", act_b="rn", wt_b="x2", wt_b3="bf16x2")
ROUND = "trunc"


def acc32(s64):
    """fp32 accumulator after adding: round toward zero (tcgen05) or to nearest (A/B)."""
    r = s64.to(torch.float32)
    if ROUND == "trunc":
        over = r.double().abs() > s64.abs()
        r = torch.where(over, torch.nextafter(r, torch.zeros_like(r)), r)
    return r


def rn32(s64):
    return s64.to(torch.float32).double()


def f16_pair(v):
    v32 = v.to(torch.float32)
    hi = q_rn(v32).to(torch.float16).to(torch.float32)
    lo = ((v32 - hi) * 4096.0).to(torch.float16).to(torch.float32) / 4096.0
    return hi.double(), lo.double()


def conv3_emulated(x, w, b, padding):
    """x [B,H,W,T,32] (fp32-representable values in fp64), w [3,3,3,32,32]"""
    xh, xl = f16_pair(x)
    wh, wl = f16_pair(w)
    if padding == "same":
        pad = (0, 0, 1, 1, 1, 1, 1, 1)
        xh, xl = torch.nn.functional.pad(xh, pad), torch.nn.functional.pad(xl, pad)
    B, H, W, T, _ = xh.shape
    Ho, Wo, To = H - 2, W - 2, T - 2
    thirds, corr = [], torch.zeros(B, Ho, Wo, To, w.shape[-1], dtype=torch.float64)
    for dw in range(3):
        acc = torch.zeros(B, Ho, Wo, To, w.shape[-1], dtype=torch.float32)
        for dt in range(3):
            for dh in range(3):
                sh = xh[:, dh:dh + Ho, dw:dw + Wo, dt:dt + To]
                sl = xl[:, dh:dh + Ho, dw:dw + Wo, dt:dt + To]
                for ks in range(2):
                    c = slice(16 * ks, 16 * ks + 16)
                    acc = acc32(acc.double() + sh[..., c] @ wh[dh, dw, dt, c])
                corr += sl @ wh[dh, dw, dt] + sh @ wl[dh, dw, dt]
        thirds.append(acc.double())
    out = rn32(rn32(thirds[1] + thirds[0]) + thirds[2])
    out = rn32(out + rn32(corr))
    return rn32(out + b)


def fused_emulated(x, We, be, Wd, bd):
    xs = x.shape
    X = x.reshape(-1, xs[-1]).to(torch.float32)
    we, wd = We.reshape(We.shape[-2], We.shape[-1]).to(torch.float32), Wd.reshape(Wd.shape[-2], Wd.shape[-1])
    xh = q_rn(X)
    xl = q_tr(X - xh)                                   # the MMA truncates the raw fp32 lo operand
    wh = q_rn(we)
    wl = q_tr(we - wh)
    xh, xl, wh, wl = xh.double(), xl.double(), wh.double(), wl.double()
    acc = torch.zeros(X.shape[0], we.shape[1], dtype=torch.float32)
    for a_, b_ in ((xh, wh), (xl, wh), (xh, wl)):
        for ks in range(4):
            c = slice(8 * ks, 8 * ks + 8)
            acc = acc32(acc.double() + a_[:, c] @ b_[c])
    beh = q_rn(be.to(torch.float32))
    bel = q_tr(be.to(torch.float32) - beh)
    acc = acc32(acc.double() + (beh.double() + bel.double()))
    E = torch.relu(acc.double())
    eh, el = f16_pair(E)
    dh_, dl_ = f16_pair(wd)
    accd = torch.zeros(X.shape[0], wd.shape[1], dtype=torch.float32)
    for ks in range(16):
        c = slice(16 * ks, 16 * ks + 16)
        accd = acc32(accd.double() + eh[:, c] @ dh_[c])
    corr = el @ dh_ + eh @ dl_
    D = rn32(rn32(accd.double() + rn32(corr)) + bd)
    return D.reshape(*xs[:-1], wd.shape[-1])


class EmulatedModel(TensorCoreModel):
    """forward VALUES from the emulated MMA chains, gradients through the exact-accumulation graph of the parent"""

    def _tc(self, p, name, x, padding, relu, store=True):
        y = super()._tc(p, name, x, padding, False, store)
        w = wn_kernel(p[name + "/v"], p[name + "/g"])
        if w.dim() == 5 and w.shape[0] == 3 and w.shape[3] == 32:
            with torch.no_grad():
                ye = conv3_emulated(x.detach(), w.detach(), p[name + "/bias"].detach(), padding)
            y = y + (ye - y).detach()
        return torch.relu(y) if relu else y

    def forward(self, p, x, return_taps=False):
        self._p = p
        return super().forward(p, x, return_taps)


def patch_blocks(model):
    """expConv_i + decConv_i as the fused kernel: wrap the two _tc calls of a block"""
    orig = model._tc

    def tc(p, name, x, padding, relu, store=True):
        if name.startswith("expConv_"):
            i = name.split("_")[1]
            model._blk = (x, i)
            return orig(p, name, x, padding, relu, store)
        if name.startswith("decConv_"):
            y = orig(p, name, x, padding, relu, store)
            x0, i = model._blk
            with torch.no_grad():
                ye = fused_emulated(x0.detach(), wn_kernel(p[f"expConv_{i}/v"], p[f"expConv_{i}/g"]).detach(), p[f"expConv_{i}/bias"].detach(),
                                    wn_kernel(p[f"decConv_{i}/v"], p[f"decConv_{i}/g"]).detach(), p[f"decConv_{i}/bias"].detach())
            return y + (ye - y).detach()
        return orig(p, name, x, padding, relu, store)

    model._tc = tc
    return model


def main():
    global ROUND
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    ROUND = sys.argv[2] if len(sys.argv) > 2 else "trunc"
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
    tm = patch_blocks(EmulatedModel(NIR[0], NIR[1], 6, **FULL, mode=MODE))
    t0 = time.time()
    loss, g, sr = run(tm, p, lr, hr, mask)
    errs = sorted((float(np.abs(g[k] - r).max() / np.abs(r).max()), k) for k, r in ref_g.items() if np.abs(r).max() > 0)
    sr_err = float(np.abs(sr - ref_sr).max() / np.abs(ref_sr).max()) if ref_sr is not None else float("nan")
    print(f"B={B} accumulation={ROUND}: loss rel {abs(loss - ref_loss) / ref_loss:.2e}, SR max rel err {sr_err:.2e}, gradients worst {errs[-1][0]:.2e} at {errs[-1][1]}, "
          f"median {errs[len(errs) // 2][0]:.2e}, over 1e-3: {sum(e[0] > 1e-3 for e in errs)}/{len(errs)}  [{time.time() - t0:.0f} s]", flush=True)


if __name__ == "__main__":
    main()
