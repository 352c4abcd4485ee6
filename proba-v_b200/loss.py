"""Losses with the reference's surface (reference models/loss.py:8-97): Losses(targetShape, cropBorder=3,
bitDepth=16) and shiftCompensated{L1Loss,L2Loss,L1EdgeLoss,cPSNR}(patchHR, maskHR, predPatchHR).
Every call is one launch of the fused shift-search kernel (csrc/shift_loss.cu) through pv_shift_loss*."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _buf
from ._lib import PV_LOSS, check, lib


class Losses:
    def __init__(self, targetShape=(96, 96, 1), cropBorder=3, bitDepth=16):
        self.targetShapeHeight, self.targetShapeWidth, self.targetShapeChannels = targetShape
        self.cropBorder = cropBorder
        self.maxPixelShift = 2 * cropBorder
        self.numBytes = 2 ** bitDepth - 1
        self.cropSizeHeight = self.targetShapeHeight - self.maxPixelShift
        self.cropSizeWidth = self.targetShapeWidth - self.maxPixelShift
        if bitDepth != 16:
            raise ValueError("only bitDepth=16 is built (the reference never passes another value)")
        if self.targetShapeChannels != 1:
            raise ValueError("only single-channel targets are built")

    # ---- one fused evaluation ------------------------------------------------------------------
    def evaluate(self, kind: str, patchHR, maskHR, predPatchHR, want_grad: bool = False, want_stack: bool = False,
                 grad_scale: float = None) -> dict:
        """All outputs of one kernel pass: loss_per_sample, best_shift, clear_count, cpsnr, mean_loss[, dsr, stack]."""
        shp = tuple(predPatchHR.shape)
        if len(shp) != 4 or shp[3] != 1 or tuple(patchHR.shape) != shp or tuple(maskHR.shape) != shp:
            raise ValueError(f"expected HR, mask, SR all [B,H,W,1]; got {tuple(patchHR.shape)}, {tuple(maskHR.shape)}, {shp}")
        B, H, W, _ = shp
        if (H, W) != (self.targetShapeHeight, self.targetShapeWidth):
            raise ValueError(f"targetShape {(self.targetShapeHeight, self.targetShapeWidth)} != data {(H, W)}")
        k = PV_LOSS[kind]
        gs = (1.0 / B) if grad_scale is None else float(grad_scale)
        S2 = (self.maxPixelShift + 1) ** 2
        if _buf.is_cuda_tensor(predPatchHR):
            import torch
            dev = predPatchHR.device
            sr = _buf.dev_tensor(predPatchHR, torch.float32, dev)
            hr = _buf.dev_tensor(patchHR, torch.float32, dev)
            mk = _buf.dev_tensor(maskHR, torch.uint8, dev)
            f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
            out = dict(loss_per_sample=f(B), cpsnr=f(B), mean_loss=f(1),
                       best_shift=torch.empty(B, dtype=torch.int32, device=dev),
                       clear_count=torch.empty(B, dtype=torch.int32, device=dev),
                       dsr=f(B, H, W, 1) if want_grad else None, stack=f(B, S2, 4) if want_stack else None)
            check(lib().pv_shift_loss(k, _buf.ptr(hr), _buf.ptr(mk), _buf.ptr(sr), B, H, W, self.cropBorder, gs,
                                      _buf.ptr(out["loss_per_sample"]), _buf.ptr(out["best_shift"]), _buf.ptr(out["clear_count"]),
                                      _buf.ptr(out["cpsnr"]), _buf.ptr(out["mean_loss"]), _buf.ptr(out["dsr"]),
                                      _buf.ptr(out["stack"]), _buf.current_stream_ptr(dev)))
            return out
        sr = _buf.host_array(predPatchHR, np.float32)
        hr = _buf.host_array(patchHR, np.float32)
        mk = _buf.host_array(maskHR, np.uint8)
        out = dict(loss_per_sample=np.empty(B, np.float32), cpsnr=np.empty(B, np.float32), mean_loss=np.empty(1, np.float32),
                   best_shift=np.empty(B, np.int32), clear_count=np.empty(B, np.int32),
                   dsr=np.empty((B, H, W, 1), np.float32) if want_grad else None,
                   stack=np.empty((B, S2, 4), np.float32) if want_stack else None)
        check(lib().pv_shift_loss_host(k, _buf.ptr(hr), _buf.ptr(mk), _buf.ptr(sr), B, H, W, self.cropBorder, gs,
                                       _buf.ptr(out["loss_per_sample"]), _buf.ptr(out["best_shift"]), _buf.ptr(out["clear_count"]),
                                       _buf.ptr(out["cpsnr"]), _buf.ptr(out["mean_loss"]), _buf.ptr(out["dsr"]), _buf.ptr(out["stack"])))
        return out

    @staticmethod
    def _scalar(x):
        return x[0] if hasattr(x, "is_cuda") else float(x[0])

    # ---- reference method names ------------------------------------------------------------------
    def shiftCompensatedL1Loss(self, patchHR, maskHR, predPatchHR):
        return self._scalar(self.evaluate("l1", patchHR, maskHR, predPatchHR)["mean_loss"])

    def shiftCompensatedL2Loss(self, patchHR, maskHR, predPatchHR):
        return self._scalar(self.evaluate("l2", patchHR, maskHR, predPatchHR)["mean_loss"])

    def shiftCompensatedL1EdgeLoss(self, patchHR, maskHR, predPatchHR):
        return self._scalar(self.evaluate("sobel_l1_mix", patchHR, maskHR, predPatchHR)["mean_loss"])

    def shiftCompensatedcPSNR(self, patchHR, maskHR, predPatchHR):
        return self.evaluate("l1", patchHR, maskHR, predPatchHR)["cpsnr"]

    def shiftCompensatedRevSSIM(self, patchHR, maskHR, predPatchHR):
        raise NotImplementedError("shiftCompensatedRevSSIM is broken in the reference (loss.py:108-109) and out of scope")


LOSS_KIND_OF_METHOD = {"shiftCompensatedL1Loss": "l1", "shiftCompensatedL2Loss": "l2",
                       "shiftCompensatedL1EdgeLoss": "sobel_l1_mix"}


def loss_from_config(loss: Losses, name: str):
    """train.py:93-100: cfg 'loss' -> bound method."""
    table = {"l1": loss.shiftCompensatedL1Loss, "sobel_l1_mix": loss.shiftCompensatedL1EdgeLoss,
             "l2": loss.shiftCompensatedL2Loss, "l1msssim": loss.shiftCompensatedRevSSIM}
    if name not in table:
        raise ValueError(f"unknown loss {name!r} (cfg [Train] loss is one of {sorted(table)})")
    return table[name]
