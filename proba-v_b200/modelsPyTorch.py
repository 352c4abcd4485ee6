"""PyTorch face of the B200 engine -- the counterpart of the reference's models/modelsPyTorch.py.

The reference file (models/modelsPyTorch.py:10-151: Conv3DResNet, ResBlockConv3D, ExpConv3D, DecConv3D, NormConv3D,
UpSampleConv3D, Reshape, DepthToSpace) is an unfinished draft of the 3D-WDSR graph in torch.nn: its blocks build fresh modules
inside forward() and reference undefined names, so it cannot be run.  What a user of that file needs is
the SAME graph as models/modelsTF.py behind torch.nn / torch.optim / autograd, and that is what this module provides:

  * `Conv3DResNet` -- torch.nn.Module with the draft's constructor signature.  Its single nn.Parameter `theta` IS the engine's
    flat weight arena (zero-copy view, TensorFlow variable order: <layer>/v, /g, /bias); `named_variables()` gives per-tensor views.
  * forward(x) takes the draft's channels-first layout [B, C=1, T, H, W] (modelsPyTorch.py:16 comment "(1, 9, 34, 34)") and returns
    [B, 1, s*P, s*P]; it runs the hand-written sm_100a kernels through the C-ABI (pv_trainer_forward) and is differentiable:
    backward hands dL/dSR to pv_trainer_backward, which fills the gradient arena that `theta.grad` then views.
  * any torch loss / optimizer works on top (`ShiftL1Loss` below wraps the fused shift-search loss kernel as an autograd Function).

There is no CPU fallback: the module lives on an sm_100 device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _buf
from ._lib import check, lib
from .loss import PV_LOSS
from .models import WDSRConv3D

NIR = (8075.2045, 3160.7272)        # train.py:47-49


class _EngineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, theta, mod):
        B = int(x_cl.shape[0])
        side = mod.model.out_side
        sr = torch.empty((B, side, side, 1), dtype=torch.float32, device=x_cl.device)
        if mod._theta_version != theta._version:         # an optimizer stepped: the effective (weight-normalised) kernels are stale
            mod.model.param_arena()
            mod._theta_version = theta._version
        check(lib().pv_trainer_forward(mod._trainer, _buf.ptr(x_cl), B, _buf.ptr(sr), _buf.current_stream_ptr(x_cl.device)))
        ctx.mod, ctx.B = mod, B
        return sr

    @staticmethod
    def backward(ctx, g):
        mod = ctx.mod
        g = g.contiguous().float()
        check(lib().pv_trainer_backward(mod._trainer, _buf.ptr(g), ctx.B, _buf.current_stream_ptr(g.device)))
        return None, mod._grad_view.clone(), None          # the LR input needs no gradient (trainClass.py:131 differentiates the weights)


class Conv3DResNet(nn.Module):
    """modelsPyTorch.py:10-37 signature; the graph is modelsTF.py:15-203 (trunk + reducers + upscale + the learned 2-D skip path).

    inputSize = (C, T, H, W) of one LR patch stack, e.g. (1, 9, 22, 22) for cfg/p16t9c85r12 (H = patch_size + max_shift)."""

    def __init__(self, inputSize: tuple, upSampleScale: int, numResBlocks: int, kernelSize: tuple, numFilters: int, expRate: int,
                 decayRate: float, maxShift: int = 6, band_stats: tuple = NIR, precision: str = "tf32x3", device: int = None, seed: int = 0):
        super().__init__()
        cin, T, H, W = inputSize
        if cin != 1 or H != W:
            raise ValueError("PROBA-V patches are single-channel and square")
        self.model = WDSRConv3D("superResolutionNet", "NIR", band_stats[0], band_stats[1], maxShift).build(
            scale=upSampleScale, numFilters=numFilters, kernelSize=tuple(kernelSize), numResBlocks=numResBlocks, expRate=expRate,
            decayRate=decayRate, numImgLR=T, patchSizeLR=H - maxShift, isGrayScale=True, precision=precision, device=device, seed=seed)
        h = C.c_void_p()
        check(lib().pv_trainer_create(self.model._h, 0, C.c_float(0.0), PV_LOSS["l1"], C.byref(h)))   # optimizer slots unused: torch.optim steps theta
        self._trainer = h
        self.theta = nn.Parameter(self.model.param_arena(), requires_grad=True)
        p, n = C.c_void_p(), C.c_int64()
        check(lib().pv_trainer_grad_arena(self._trainer, C.byref(p), C.byref(n)))
        self._grad_view = _buf.view_device_floats(p.value, n.value, f"cuda:{self.model.device}")
        self._theta_version = self.theta._version

    def named_variables(self):
        """(name, view) per TensorFlow variable (<layer>/v | /g | /bias), views into `theta`."""
        for v in self.model.trainable_variables:
            yield v.name, self.theta.data[v.offset:v.offset + v.numel].view(*v.shape)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected [B, 1, T, H, W], got {tuple(x.shape)}")
        x_cl = x.to(self.theta.device, torch.float32).permute(0, 3, 4, 2, 1).contiguous()        # -> [B, H, W, T, 1] (modelsTF.py:19)
        sr = _EngineFn.apply(x_cl, self.theta, self)
        return sr.permute(0, 3, 1, 2)                                                            # [B, 1, sP, sP]

    def __del__(self):
        try:
            if getattr(self, "_trainer", None):
                lib().pv_trainer_destroy(self._trainer)
                self._trainer = None
        except Exception:
            pass


class _ShiftLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sr, hr, mask, kind):
        B, H, W = int(sr.shape[0]), int(sr.shape[1]), int(sr.shape[2])
        dev = sr.device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)           # noqa: E731
        loss_ps, cps, mean, dsr = f(B), f(B), f(1), f(B, H, W, 1)
        best = torch.empty(B, dtype=torch.int32, device=dev)
        cnt = torch.empty(B, dtype=torch.int32, device=dev)
        check(lib().pv_shift_loss(PV_LOSS[kind], _buf.ptr(hr), _buf.ptr(mask), _buf.ptr(sr), B, H, W, 3, 1.0 / B, _buf.ptr(loss_ps),
                                  _buf.ptr(best), _buf.ptr(cnt), _buf.ptr(cps), _buf.ptr(mean), _buf.ptr(dsr), None,
                                  _buf.current_stream_ptr(dev)))
        ctx.save_for_backward(dsr)
        return mean[0]

    @staticmethod
    def backward(ctx, g):
        (dsr,) = ctx.saved_tensors
        return dsr * g, None, None, None


class ShiftL1Loss(nn.Module):
    """Losses.shiftCompensatedL1Loss (loss.py:73-84) as a torch criterion on channels-first tensors [B, 1, H, W]; `kind` may also be
    "l2" or "sobel_l1_mix".  Forward and the closed-form backward both come from the one fused CUDA kernel."""

    def __init__(self, kind: str = "l1"):
        super().__init__()
        self.kind = kind

    def forward(self, sr: torch.Tensor, hr: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        cl = lambda t, dt: t.permute(0, 2, 3, 1).contiguous().to(dt)              # noqa: E731
        return _ShiftLossFn.apply(cl(sr, torch.float32), cl(hr.to(sr.device), torch.float32), cl(mask.to(sr.device), torch.uint8), self.kind)
