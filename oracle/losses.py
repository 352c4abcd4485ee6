"""Oracle: shift-compensated losses (CPU, torch; differentiable).  TEST INFRASTRUCTURE ONLY.

Restates reference models/loss.py (class Losses) line by line:
  __init__ :13-35 | cPSNR :37-53 | L2 :55-71 | L1 :73-84 | L1Edge :86-97
  stack*   :126-180 | computeBiasBrightness :182-187 | compute* :219-238
and utils/utils.py:42-44 (cropImage).  shiftCompensatedRevSSIM (:99-124,:189-217) is out of
scope (broken as written; SURVEY §2.1 #2b).
Reference quirks reproduced on purpose:
  * only the SR is multiplied by the mask; HR enters un-masked (loss.py:141-152);
  * bias b = (sum(HR) - sum(SR*mask)) / N with sum(HR) over ALL window pixels (:182-187);
  * N = 0 is unguarded (inf/NaN).
tf.reduce_min/max tie-splitting is NOT modelled here: torch.min/max autograd routes the gradient
to one arg-extremum; the CUDA kernel uses first-minimum (argmin) -- documented in DESIGN.md.
PARITY UNPINNED (no TensorFlow in the image).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch
import torch.nn.functional as F


def cropImage(img: torch.Tensor, h0: int, lh: int, w0: int, lw: int, dtype) -> torch.Tensor:
    """utils/utils.py:42-44 -- slice then cast to float."""
    return img[:, h0:h0 + lh, w0:w0 + lw, :].to(dtype)


def sobel_edges(img: torch.Tensor) -> torch.Tensor:
    """tf.image.sobel_edges on [B,H,W,C] -> [B,H,W,C,2] (dy, dx); REFLECT pad 1 (Appendix B.5)."""
    B, H, W, C = img.shape
    ky = torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]], dtype=img.dtype)
    kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]], dtype=img.dtype)
    x = img.permute(0, 3, 1, 2).reshape(B * C, 1, H, W)
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    k = torch.stack([ky, kx]).unsqueeze(1)                 # [2,1,3,3]
    y = F.conv2d(x, k)                                     # [B*C,2,H,W]
    return y.reshape(B, C, 2, H, W).permute(0, 3, 4, 1, 2)


class OracleLosses:
    def __init__(self, targetShape=(96, 96, 1), cropBorder=3, bitDepth=16, dtype=torch.float64):
        self.H, self.W, self.C = targetShape
        self.cropBorder = cropBorder
        self.maxPixelShift = 2 * cropBorder                 # loss.py:18
        self.numBytes = 2 ** bitDepth - 1                   # :19
        self.pi = 0.7                                       # :21
        self.cropH = self.H - self.maxPixelShift            # :23
        self.cropW = self.W - self.maxPixelShift            # :24
        self.dtype = dtype

    # ---- per-shift pieces --------------------------------------------------------------
    def _prep(self, i, j, patchHR, maskHR, cropPred):
        h = cropImage(patchHR, i, self.cropH, j, self.cropW, self.dtype)    # :141
        m = cropImage(maskHR, i, self.cropH, j, self.cropW, self.dtype)     # :142
        predM = cropPred * m                                                # :143
        N = m.sum(dim=(1, 2, 3))                                            # :144
        b = (1.0 / N) * (h - predM).sum(dim=(1, 2, 3))                      # :184
        b = b.reshape(-1, 1, 1, 1)                                          # :186
        corrM = (cropPred + b) * m                                          # :148-149
        return h, m, N, b.reshape(-1), corrM

    def _score(self, kind, N, h, corrM):
        if kind == "l1":
            return (1.0 / N) * (h - corrM).abs().sum(dim=(1, 2, 3))         # :226-228
        if kind == "l2":
            return (1.0 / N) * ((h - corrM) ** 2).sum(dim=(1, 2, 3))        # :230-232
        if kind == "cpsnr":
            l2 = (1.0 / N) * ((h - corrM) ** 2).sum(dim=(1, 2, 3))
            return 10.0 * (torch.log(self.numBytes ** 2 / l2) / math.log(10.0))   # :234-238
        if kind == "l1edge":
            l1 = (1.0 / N) * (h - corrM).abs().sum(dim=(1, 2, 3))
            sob = (1.0 / N) * (sobel_edges(h) - sobel_edges(corrM)).abs().sum(dim=(1, 2, 3, 4))
            return self.pi * l1 + (1 - self.pi) * sob                       # :219-224
        raise ValueError(kind)

    def stack(self, kind, patchHR, maskHR, predPatchHR) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Scores [S*S, B] in reference stack order (index i*(S)+j), clear counts [S*S,B], bias [S*S,B]."""
        cropPred = cropImage(predPatchHR, self.cropBorder, self.cropH, self.cropBorder, self.cropW, self.dtype)
        scores, counts, biases = [], [], []
        for i in range(self.maxPixelShift + 1):
            for j in range(self.maxPixelShift + 1):
                h, m, N, b, corrM = self._prep(i, j, patchHR, maskHR, cropPred)
                scores.append(self._score(kind, N, h, corrM))
                counts.append(N)
                biases.append(b)
        return torch.stack(scores), torch.stack(counts), torch.stack(biases)

    # ---- public API, same names as the reference ---------------------------------------
    def shiftCompensatedL1Loss(self, patchHR, maskHR, predPatchHR):
        s, _, _ = self.stack("l1", patchHR, maskHR, predPatchHR)
        return s.min(dim=0).values.mean()                                   # :82-84

    def shiftCompensatedL2Loss(self, patchHR, maskHR, predPatchHR):
        s, _, _ = self.stack("l2", patchHR, maskHR, predPatchHR)
        return s.min(dim=0).values.mean()                                   # :69-71

    def shiftCompensatedL1EdgeLoss(self, patchHR, maskHR, predPatchHR):
        s, _, _ = self.stack("l1edge", patchHR, maskHR, predPatchHR)
        return s.min(dim=0).values.mean()                                   # :95-97

    def shiftCompensatedcPSNR(self, patchHR, maskHR, predPatchHR):
        s, _, _ = self.stack("cpsnr", patchHR, maskHR, predPatchHR)
        return s.max(dim=0).values                                          # :51-53  ([B], un-reduced)

    # ---- extras the parity tests need ----------------------------------------------------
    def details(self, kind, patchHR, maskHR, predPatchHR):
        """per-sample (best score, best shift index (first extremum), clear count at that shift)."""
        s, n, _ = self.stack(kind, patchHR, maskHR, predPatchHR)
        if kind == "cpsnr":
            idx = s.argmax(dim=0)
        else:
            idx = s.argmin(dim=0)
        best = s.gather(0, idx[None])[0]
        cnt = n.gather(0, idx[None])[0]
        return best, idx, cnt, s

    def l1_grad_closed_form(self, patchHR, maskHR, predPatchHR):
        """d mean_b(min_shift L1) / d pred, evaluated at the first arg-min shift."""
        B = patchHR.shape[0]
        _, idx, _, _ = self.details("l1", patchHR, maskHR, predPatchHR)
        cropPred = cropImage(predPatchHR, self.cropBorder, self.cropH, self.cropBorder, self.cropW, self.dtype)
        g = torch.zeros_like(predPatchHR, dtype=self.dtype)
        S = self.maxPixelShift + 1
        cb = self.cropBorder
        for b in range(B):
            i, j = int(idx[b]) // S, int(idx[b]) % S
            h = patchHR[b:b + 1, i:i + self.cropH, j:j + self.cropW, :].to(self.dtype)
            m = maskHR[b:b + 1, i:i + self.cropH, j:j + self.cropW, :].to(self.dtype)
            p = cropPred[b:b + 1]
            N = m.sum()
            bias = (h - p * m).sum() / N
            r = h - (p + bias) * m
            s = torch.sign(r)
            g[b:b + 1, cb:cb + self.cropH, cb:cb + self.cropW, :] = (m / N) * (-s + (s * m).sum() / N) / B
        return g
