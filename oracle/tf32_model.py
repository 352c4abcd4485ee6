"""Oracle-side NUMERICS MODEL of the tensor-core engine (CPU, torch fp64 accumulate).  TEST / STUDY INFRASTRUCTURE ONLY.

The same graph as oracle/wdsr.py (reference models/modelsTF.py:7-203) and the same autograd as trainClass.py:126-131,
but every convolution that the row engine runs on the tensor cores takes its operands through a quantiser, in the
forward pass AND in both backward products, exactly where the CUDA kernels round:

    forward      y  = conv(q_a(x), q_w(w)) + b           stored activations are rounded when they only feed MMAs
    data grad    gx = conv^T(q_g(gy), q_w(w))
    weight grad  gw = corr(q_a(x), q_g(gy)),  gb = sum(gy)

`q_*` is one of
    "rn"    round-to-nearest-away to tf32 (10 explicit mantissa bits)   -- cvt.rna.tf32.f32 / the half-ulp bump
    "tr"    truncation to tf32                                           -- what tcgen05 kind::tf32 does to a raw fp32 operand
    "x2"    hi + lo split, hi = tr(x), lo = tr(x - hi): two MMAs per product, ~21 mantissa bits
    "f16p"  the packed fp16 pair row of the final engine (hi = fp16(tf32(x)), lo scaled by 2^12): ~22 bits;  "bf16x2" the bf16 pair (16 bits)
    "none"  exact (fp32 CUDA-core engine)
`act_b` / `wt_b` (default: the same as act / wt) are the quantisers the BACKWARD products apply to the saved activation /
the weights; `stream` says whether the residual stream A_i (forward) and its gradient (backward) are rounded at every block boundary
(the engine stores them rounded) or kept in fp32.

It exists to answer, without GPU time, "which roundings put the gradients outside the north_star's 1e-3 bar, and what is
the cheapest set of compensated (split) operands that brings them back in" -- scripts/tf32_study.py runs the grid.
Nothing in the product imports this file.
"""
from __future__ import annotations

from typing import Dict

import torch

from .wdsr import OracleWDSR, conv_cl, depth_to_space, reducer_pads, reducer_plan, reflect_pad_hwt, wn_kernel


def _bits(x32: torch.Tensor) -> torch.Tensor:
    return x32.contiguous().view(torch.int32)


def q_rn(x: torch.Tensor) -> torch.Tensor:
    b = _bits(x.to(torch.float32))
    b = (b + 0x1000) & ~0x1FFF
    return b.view(torch.float32).to(x.dtype)


def q_tr(x: torch.Tensor) -> torch.Tensor:
    b = _bits(x.to(torch.float32)) & ~0x1FFF
    return b.view(torch.float32).to(x.dtype)


def q_x2(x: torch.Tensor) -> torch.Tensor:
    x32 = x.to(torch.float32)
    hi = q_tr(x32)
    lo = q_tr(x32 - hi)
    return (hi.double() + lo.double()).to(x.dtype)


def q_bf16x2(x: torch.Tensor) -> torch.Tensor:
    """hi + lo bf16 split (round-to-nearest each): ~16 mantissa bits, two (three) bf16 MMAs per product."""
    x32 = x.to(torch.float32)
    hi = x32.to(torch.bfloat16).to(torch.float32)
    lo = (x32 - hi).to(torch.bfloat16).to(torch.float32)
    return (hi.double() + lo.double()).to(x.dtype)


def q_bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)


def q_f16p(x: torch.Tensor) -> torch.Tensor:
    """the packed fp16 pair row of csrc/rows.h: fp16(tf32(x)) + fp16(2^12 (x - fp16(tf32(x)))) / 2^12 (~22 mantissa bits for O(1) values)."""
    x32 = x.to(torch.float32)
    hi = q_rn(x32).to(torch.float16).to(torch.float32)
    lo = ((x32 - hi) * 4096.0).to(torch.float16).to(torch.float32) / 4096.0
    return (hi.double() + lo.double()).to(x.dtype)


Q = {"rn": q_rn, "tr": q_tr, "x2": q_x2, "none": lambda x: x, "bf16": q_bf16, "bf16x2": q_bf16x2, "f16p": q_f16p}


class _QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, padding, qa, qw, qg, qab, qwb):
        xq, wq = Q[qa](x), Q[qw](w)
        # the backward products may see the operands through a different quantiser than the forward product
        # (a compensated forward stores un-rounded fp32; a single-pass backward MMA then truncates it)
        ctx.save_for_backward(xq if qab == qa else Q[qab](x), wq if qwb == qw else Q[qwb](w))
        ctx.padding, ctx.qg = padding, qg
        return conv_cl(xq, wq, b, padding, False)

    @staticmethod
    def backward(ctx, gy):
        xq, wq = ctx.saved_tensors
        gq = Q[ctx.qg](gy)
        with torch.enable_grad():
            xl, wl = xq.detach().requires_grad_(True), wq.detach().requires_grad_(True)
            y = conv_cl(xl, wl, None, ctx.padding, False)
            gx, gw = torch.autograd.grad(y, (xl, wl), gq)
        gb = gy.sum(dim=tuple(range(gy.dim() - 1)))
        return gx, gw, gb, None, None, None, None, None, None


class _QGrad(torch.autograd.Function):
    """identity whose backward quantises the incoming gradient (a stored, rounded gradient tensor)"""
    @staticmethod
    def forward(ctx, x, q):
        ctx.q = q
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return Q[ctx.q](g), None


class _QStraight(torch.autograd.Function):
    """forward quantisation with a straight-through gradient (a stored, rounded activation)"""
    @staticmethod
    def forward(ctx, x, q):
        return Q[q](x)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _FusedExpDec(torch.autograd.Function):
    """expConv (1x1x1) -> ReLU -> decConv (1x1x1) of one block as the engine's three fused kernels compute it.

    forward        E = relu(q_a(x) q_w(We) + be),  D = q_a(E) q_w(Wd) + bd                       (resfront forward)
    data gradient  gZ = (q_g(gD) q_wb(Wd)^T) . mask,  gx = q_g(gZ) q_wb(We)^T                     (resfront_bwd_data)
    weight kernel  RECOMPUTES  E_w = relu(q_xw(x) q_ww(We) + be)  and  gZ_w = (q_g(gD) q_ww(Wd)^T) . mask  on chip, then
                   dWd = q_rn(E_w)^T q_g(gD),  dWe = q_xw(x)^T q_rn(gZ_w)                         (resfront_bwd_weight)
    `ww` / `xw` are the quantisers of that recomputation (the kernel has single-pass tf32 operands: "rn")."""
    @staticmethod
    def forward(ctx, x, We, be, Wd, bd, m):
        xs = x.shape
        X = x.reshape(-1, xs[-1])
        we, wd = We.reshape(We.shape[-2], We.shape[-1]), Wd.reshape(Wd.shape[-2], Wd.shape[-1])
        E = torch.relu(Q[m["act"]](X) @ Q[m["wt"]](we) + be)
        D = Q[m["act"]](E) @ Q[m["wt"]](wd) + bd
        ctx.save_for_backward(X, we, be, wd, E)
        ctx.m, ctx.xs, ctx.ws = m, xs, (We.shape, Wd.shape)
        return D.reshape(*xs[:-1], wd.shape[-1])

    @staticmethod
    def backward(ctx, gD):
        X, we, be, wd, E = ctx.saved_tensors
        m = ctx.m
        G = Q[m["grad"]](gD.reshape(-1, gD.shape[-1]))
        mask = (E > 0).to(G.dtype)
        wtb = m.get("wt_b") or m["wt"]
        gZ = (G @ Q[wtb](wd).t()) * mask
        gx = Q[m["grad"]](gZ) @ Q[wtb](we).t()
        ww, xw = m.get("ww", "rn"), m.get("xw", "rn")
        Xw = Q[xw](X)
        Ew = torch.relu(Xw @ Q[ww](we) + be)
        gZw = (G @ Q[ww](wd).t()) * mask
        dWd = q_rn(Ew).t() @ G
        dWe = Xw.t() @ q_rn(gZw)
        return gx.reshape(ctx.xs), dWe.reshape(ctx.ws[0]), gZw.sum(0), dWd.reshape(ctx.ws[1]), gD.reshape(-1, gD.shape[-1]).sum(0), None


DEFAULT = dict(act="rn", wt="rn", grad="rn", stream="rn", gstream="rn")      # the round-1 tf32 engine


class TensorCoreModel(OracleWDSR):
    """OracleWDSR with the tensor-core engine's operand quantisation; `mode` overrides DEFAULT."""

    def __init__(self, *a, mode=None, **k):
        super().__init__(*a, **k)
        self.mode = dict(DEFAULT)
        self.mode.update(mode or {})

    def _tc(self, p, name, x, padding, relu, store=True):
        m = self.mode
        w = wn_kernel(p[name + "/v"], p[name + "/g"])
        wtb = m.get("wt_b") or m["wt"]
        if w.dim() == 5 and w.shape[0] == 3 and m.get("wt_b3"):     # 3x3x3 layers may use a different backward weight quantiser
            wtb = m["wt_b3"]
        y = _QConv.apply(x, w, p[name + "/bias"], padding, m["act"], m["wt"], m["grad"], m.get("act_b") or m["act"], wtb)
        return torch.relu(y) if relu else y

    def forward(self, p: Dict[str, torch.Tensor], x: torch.Tensor, return_taps: bool = False):
        m = self.mode
        meanLR = x.mean(dim=3)
        xn = (x - self.mean) / self.std
        mn = (meanLR - self.mean) / self.std
        h = self._wn(p, "mainConv1", xn, "same", True)        # Cin = 1: CUDA cores, fp32
        h = _QStraight.apply(_QGrad.apply(h, m["gstream"]), m["stream"])
        for i in range(self.numResBlocks):
            if m.get("fused"):
                d = _FusedExpDec.apply(h, wn_kernel(p[f"expConv_{i}/v"], p[f"expConv_{i}/g"]), p[f"expConv_{i}/bias"],
                                       wn_kernel(p[f"decConv_{i}/v"], p[f"decConv_{i}/g"]), p[f"decConv_{i}/bias"], m)
            else:
                e = self._tc(p, f"expConv_{i}", h, "same", True)
                d = self._tc(p, f"decConv_{i}", e, "same", False)
            n = self._tc(p, f"normConv_{i}", d, "same", False)
            h = n + h
            h = _QStraight.apply(_QGrad.apply(h, m["gstream"]), m["stream"])
        pads = reducer_pads(self.numImgLR)
        for i, (name, _k) in enumerate(reducer_plan(self.numImgLR, self.scale, self.kernelSize), start=1):
            if i in pads:
                h = reflect_pad_hwt(h, *pads[i])
            h = self._tc(p, name, h, "valid", True)
        h = self._tc(p, "upscaleConv1", h, "valid", False)
        B = x.shape[0]
        main = depth_to_space(h.reshape(B, self.patchSizeLR, self.patchSizeLR, self.scale * self.scale), self.scale)
        r = mn
        for i in range(self.scale):                           # 2-D skip path: CUDA cores, fp32
            r = self._wn(p, f"residConv{i+1}", r, "valid", i == 0)
        out = (main + depth_to_space(r, self.scale)) * self.std + self.mean
        return (out, {}) if return_taps else out
