"""Oracle: 3D-WDSR forward graph (CPU, torch).  TEST INFRASTRUCTURE ONLY.

Restates reference models/modelsTF.py:7-203 (class WDSRConv3D) with the
third-party semantics of SURVEY.md Appendix B:
  * TFA WeightNormalization(data_init=False): w = v * g * rsqrt(max(sum v^2, 1e-12)),
    norm over every axis but the last (Cout)                  (modelsTF.py:191-197)
  * Keras Conv3D/Conv2D channels-last, cross-correlation, 'same' = zero pad (B.2)
  * tf.pad(mode='reflect')                                     (modelsTF.py:157-158)
  * tf.nn.depth_to_space, NHWC "DCR" ordering                  (modelsTF.py:52,73)
All tensors are channels-last exactly as in the reference:
  LR  [B, H, W, T, 1]   ->   SR [B, scale*patch, scale*patch, 1].
PARITY UNPINNED (no TensorFlow in the image) -- see oracle/__init__.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- graph spec
def layer_specs(scale: int, numFilters: int, kernelSize, numResBlocks: int, expRate: int,
                decayRate: float, numImgLR: int, isGrayScale: bool = True) -> List[dict]:
    """Ordered list of weight-normalised conv layers, in Keras layer-creation order.

    Each spec: name, kind ('3d'|'2d'), k (kernel tuple), cin, cout.
    Follows modelsTF.py:45-53 (2D path), :55-74 (3D trunk), :76-175 (reducers), :177-189 (block).
    """
    k3 = tuple(kernelSize)
    k2 = tuple(kernelSize[:-1])
    cin0 = 1 if isGrayScale else 3
    specs = [dict(name="mainConv1", kind="3d", k=k3, cin=cin0, cout=numFilters)]
    dec = int(numFilters * decayRate)                      # modelsTF.py:182
    for i in range(numResBlocks):
        specs.append(dict(name=f"expConv_{i}", kind="3d", k=(1, 1, 1), cin=numFilters, cout=numFilters * expRate))
        specs.append(dict(name=f"decConv_{i}", kind="3d", k=(1, 1, 1), cin=numFilters * expRate, cout=dec))
        specs.append(dict(name=f"normConv_{i}", kind="3d", k=k3, cin=dec, cout=numFilters))
    for (name, kk) in reducer_plan(numImgLR, scale, k3):
        specs.append(dict(name=name, kind="3d", k=kk, cin=numFilters, cout=numFilters))
    specs.append(dict(name="upscaleConv1", kind="3d", k=k3, cin=numFilters, cout=scale * scale))
    c = cin0
    for i in range(scale):                                  # modelsTF.py:47-50
        specs.append(dict(name=f"residConv{i+1}", kind="2d", k=k2, cin=c, cout=scale * scale))
        c = scale * scale
    return specs


def reducer_plan(numImgLR: int, scale: int, k3) -> List[Tuple[str, tuple]]:
    """(layer name, kernel) of the convReducer_* layers for a given T (modelsTF.py:62-69)."""
    if numImgLR == 7 or numImgLR == 9:
        n = numImgLR // scale                               # :154, :169
    elif numImgLR == 13:
        n = 5                                               # :123-150
    elif numImgLR == 19:
        n = 10                                              # :76-121
    else:
        raise ValueError(f"num_low_res_imgs={numImgLR}: the reference graph has no reducer for it "
                         "(modelsTF.py:62-69 handles 7, 9, 13, 19)")
    out = []
    for i in range(n):
        kk = (5, 5, 5) if (numImgLR == 19 and i == 0) else tuple(k3)
        out.append((f"convReducer_{i+1}", kk))
    return out


def reducer_pads(numImgLR: int) -> Dict[int, Tuple[int, int, int]]:
    """Reflect pad (H, W, T) applied BEFORE reducer i (1-based), per reference variant."""
    if numImgLR == 9:
        return {1: (1, 1, 0)}                               # :156-158
    if numImgLR == 7:
        return {}                                           # :166-175
    if numImgLR == 13:
        return {1: (1, 1, 0), 2: (1, 1, 0), 3: (1, 1, 0)}   # :126-138
    if numImgLR == 19:
        return {1: (2, 2, 2), 2: (2, 2, 1), 3: (2, 2, 0), 4: (2, 2, 0), 5: (1, 1, 0)}   # :78-101
    raise ValueError(numImgLR)


def init_params(specs: List[dict], seed: int = 0, dtype=torch.float64, g_jitter: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights: Glorot-uniform v, g = ||v||*U(0.5,1.5), small bias.

    Names follow the TF checkpoint: '<layer>/v', '<layer>/g', '<layer>/bias'.
    """
    gen = torch.Generator().manual_seed(seed)
    p = {}
    for s in specs:
        k = s["k"]
        shape = (*k, s["cin"], s["cout"])
        rf = math.prod(k)
        limit = math.sqrt(6.0 / (rf * s["cin"] + rf * s["cout"]))
        v = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit
        nrm = torch.sqrt((v * v).reshape(-1, s["cout"]).sum(0))
        g = nrm * (0.5 + torch.rand(s["cout"], generator=gen, dtype=torch.float64)) if g_jitter else nrm
        b = (torch.rand(s["cout"], generator=gen, dtype=torch.float64) * 2 - 1) * 0.05
        p[s["name"] + "/v"] = v.to(dtype)
        p[s["name"] + "/g"] = g.to(dtype)
        p[s["name"] + "/bias"] = b.to(dtype)
    return p


# --------------------------------------------------------------------------- primitives
def wn_kernel(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """TFA WeightNormalization kernel: l2_normalize(v, axes=all but last) * g  (Appendix B.1)."""
    axes = tuple(range(v.dim() - 1))
    ss = (v * v).sum(dim=axes)
    return v * (g * torch.rsqrt(torch.clamp(ss, min=1e-12)))


def conv_cl(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, padding: str, relu: bool) -> torch.Tensor:
    """Keras ConvND, channels-last, stride 1.  x [B, *spatial, Cin]; w [*k, Cin, Cout]."""
    nd = x.dim() - 2
    perm_in = (0, nd + 1, *range(1, nd + 1))
    xt = x.permute(*perm_in)
    wt = w.permute(nd + 1, nd, *range(nd))
    pad = tuple(kk // 2 for kk in w.shape[:nd]) if padding == "same" else 0
    fn = F.conv3d if nd == 3 else F.conv2d
    y = fn(xt, wt, b, padding=pad)
    y = y.permute(0, *range(2, nd + 2), 1)
    return torch.relu(y) if relu else y


def reflect_pad_hwt(x: torch.Tensor, ph: int, pw: int, pt: int) -> torch.Tensor:
    """tf.pad(x, [[0,0],[ph,ph],[pw,pw],[pt,pt],[0,0]], mode='reflect') on [B,H,W,T,C]."""
    def refl_idx(n, p):
        idx = list(range(p, 0, -1)) + list(range(n)) + list(range(n - 2, n - 2 - p, -1))
        return torch.tensor(idx, dtype=torch.long)
    if ph:
        x = x.index_select(1, refl_idx(x.shape[1], ph))
    if pw:
        x = x.index_select(2, refl_idx(x.shape[2], pw))
    if pt:
        x = x.index_select(3, refl_idx(x.shape[3], pt))
    return x


def depth_to_space(x: torch.Tensor, bs: int) -> torch.Tensor:
    """tf.nn.depth_to_space NHWC: out[b, h*bs+i, w*bs+j, c] = in[b, h, w, (i*bs+j)*Co + c]  (B.3)."""
    B, H, W, C = x.shape
    co = C // (bs * bs)
    x = x.reshape(B, H, W, bs, bs, co).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H * bs, W * bs, co)


# --------------------------------------------------------------------------- forward
class OracleWDSR:
    """Functional restatement of WDSRConv3D(name, band, mean, std, maxShift).build(...)."""

    def __init__(self, mean: float, std: float, maxShift: int, scale: int, numFilters: int, kernelSize,
                 numResBlocks: int, expRate: int, decayRate: float, numImgLR: int, patchSizeLR: int,
                 isGrayScale: bool = True):
        self.mean, self.std, self.maxShift = float(mean), float(std), int(maxShift)
        self.scale, self.numFilters, self.kernelSize = scale, numFilters, tuple(kernelSize)
        self.numResBlocks, self.expRate, self.decayRate = numResBlocks, expRate, decayRate
        self.numImgLR, self.patchSizeLR, self.isGrayScale = numImgLR, patchSizeLR, isGrayScale
        self.specs = layer_specs(scale, numFilters, kernelSize, numResBlocks, expRate, decayRate, numImgLR, isGrayScale)
        self.in_side = patchSizeLR + maxShift               # modelsTF.py:19

    def _wn(self, p, name, x, padding, relu):
        w = wn_kernel(p[name + "/v"], p[name + "/g"])
        return conv_cl(x, w, p[name + "/bias"], padding, relu)

    def forward(self, p: Dict[str, torch.Tensor], x: torch.Tensor, return_taps: bool = False):
        """x [B, S, S, T, 1] raw DN values -> SR [B, scale*patch, scale*patch, 1] raw DN values."""
        taps = {}
        meanLR = x.mean(dim=3)                              # modelsTF.py:23  [B,H,W,1]
        xn = (x - self.mean) / self.std                     # :26, :199-200
        mn = (meanLR - self.mean) / self.std                # :27
        # ---- 3D trunk (:55-74)
        h = self._wn(p, "mainConv1", xn, "same", True)
        taps["mainConv1"] = h
        for i in range(self.numResBlocks):                  # :177-189
            e = self._wn(p, f"expConv_{i}", h, "same", True)
            d = self._wn(p, f"decConv_{i}", e, "same", False)
            n = self._wn(p, f"normConv_{i}", d, "same", False)
            h = n + h
            taps[f"block_{i}"] = h
        pads = reducer_pads(self.numImgLR)
        for i, (name, _k) in enumerate(reducer_plan(self.numImgLR, self.scale, self.kernelSize), start=1):
            if i in pads:
                h = reflect_pad_hwt(h, *pads[i])
            h = self._wn(p, name, h, "valid", True)
            taps[name] = h
        h = self._wn(p, "upscaleConv1", h, "valid", False)
        B = x.shape[0]
        main = h.reshape(B, self.patchSizeLR, self.patchSizeLR, self.scale * self.scale)   # :71
        taps["upscale"] = main
        main = depth_to_space(main, self.scale)             # :73
        # ---- 2D low-frequency path (:45-53)
        r = mn
        for i in range(self.scale):
            r = self._wn(p, f"residConv{i+1}", r, "valid", i == 0)
        taps["resid"] = r
        resid = depth_to_space(r, self.scale)
        out = (main + resid) * self.std + self.mean         # :38-41, :202-203
        return (out, taps) if return_taps else out

    def n_params(self) -> int:
        return sum(math.prod(s["k"]) * s["cin"] * s["cout"] + 2 * s["cout"] for s in self.specs)

    def macs_per_patch(self) -> int:
        """Forward multiply-accumulates per patch (SURVEY Appendix A check)."""
        S, T = self.in_side, self.numImgLR
        total = 0
        dims = {"h": S, "w": S, "t": T}
        vox = S * S * T
        nf, dec = self.numFilters, int(self.numFilters * self.decayRate)
        cin0 = 1 if self.isGrayScale else 3
        total += vox * math.prod(self.kernelSize) * cin0 * nf
        total += self.numResBlocks * vox * (nf * nf * self.expRate + nf * self.expRate * dec
                                            + math.prod(self.kernelSize) * dec * nf)
        H = W = S
        pads = reducer_pads(self.numImgLR)
        for i, (_n, k) in enumerate(reducer_plan(self.numImgLR, self.scale, self.kernelSize), start=1):
            ph, pw, pt = pads.get(i, (0, 0, 0))
            H, W, T = H + 2 * ph - (k[0] - 1), W + 2 * pw - (k[1] - 1), T + 2 * pt - (k[2] - 1)
            total += H * W * T * math.prod(k) * nf * nf
        k = self.kernelSize
        H, W, T = H - (k[0] - 1), W - (k[1] - 1), T - (k[2] - 1)
        total += H * W * T * math.prod(k) * nf * self.scale ** 2
        s2, c, h2 = self.scale ** 2, cin0, S
        for _ in range(self.scale):
            h2 -= k[0] - 1
            total += h2 * h2 * k[0] * k[1] * c * s2
            c = s2
        return total
