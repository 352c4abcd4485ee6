"""Oracle: train / test step and predict post-processing (CPU, torch).  TEST INFRASTRUCTURE ONLY.

Restates reference
  models/trainClass.py:124-135 (trainStep: fwd -> loss -> tape.gradient -> apply_gradients -> metric)
  models/trainClass.py:137-143 (testStep)
  test.py:114-160 (resolve: clip[0, 2**16] -> round-half-even; resolveByBatch; reconstruct_from_patches)
  models/testClass.py:24-39 (Enhancer.enhancePatch / reconstruct)
  utils/dataGenerator.py:108-121 (scene -> 64 patches: reflect-pad 3, window 22, stride 16)
PARITY UNPINNED (no TensorFlow in the image).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .losses import OracleLosses
from .wdsr import OracleWDSR

LOSS_METHOD = {"l1": "shiftCompensatedL1Loss", "l2": "shiftCompensatedL2Loss",
               "sobel_l1_mix": "shiftCompensatedL1EdgeLoss"}


def loss_and_grads(model: OracleWDSR, losses: OracleLosses, params: Dict[str, torch.Tensor],
                   lr, hr, mask, loss_kind="l1"):
    """Returns (loss scalar, grads dict, sr, cpsnr[B]) -- trainClass.py:126-133."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    sr = model.forward(leaf, lr)
    loss = getattr(losses, LOSS_METHOD[loss_kind])(hr, mask, sr)
    grads = torch.autograd.grad(loss, list(leaf.values()), allow_unused=True)
    g = {k: (gi if gi is not None else torch.zeros_like(v)) for (k, v), gi in zip(leaf.items(), grads)}
    with torch.no_grad():
        cpsnr = losses.shiftCompensatedcPSNR(hr, mask, sr.detach())
    return loss.detach(), g, sr.detach(), cpsnr


def train_step(model, losses, optimizer, params, lr, hr, mask, loss_kind="l1"):
    loss, g, sr, cpsnr = loss_and_grads(model, losses, params, lr, hr, mask, loss_kind)
    params = optimizer.apply_gradients({k: v.detach() for k, v in params.items()}, g)
    return params, loss, cpsnr, g


def test_step(model, losses, params, lr, hr, mask, loss_kind="l1"):
    with torch.no_grad():
        sr = model.forward(params, lr)
        loss = getattr(losses, LOSS_METHOD[loss_kind])(hr, mask, sr)
        cpsnr = losses.shiftCompensatedcPSNR(hr, mask, sr)
    return loss, cpsnr, sr


# ------------------------------------------------------------------------------- predict
def resolve(model: OracleWDSR, params, lr_batch: torch.Tensor) -> np.ndarray:
    """test.py:114-122: model -> clip_by_value(0, 2**16) -> round (half to even) -> numpy."""
    with torch.no_grad():
        sr = model.forward(params, lr_batch)
        sr = torch.clamp(sr, 0, 2 ** 16)
        sr = torch.round(sr)
    return sr.to(torch.float32).numpy()


def resolveByBatch(model, params, lr_batch, batch_size=16) -> np.ndarray:
    """test.py:125-134."""
    out = [resolve(model, params, lr_batch[s:s + batch_size]) for s in range(0, lr_batch.shape[0], batch_size)]
    return np.concatenate(out)


def reconstruct_from_patches(images: np.ndarray) -> np.ndarray:
    """test.py:149-160: n x n row-major stitch of [n*n, P, P, 1] -> [n*P, n*P, 1]."""
    n = int(len(images) ** 0.5)
    P = images.shape[1]
    rec = np.zeros((n * P, n * P, 1))
    k = 0
    for i in range(n):
        for j in range(n):
            rec[i * P:(i + 1) * P, j * P:(j + 1) * P] = images[k]
            k += 1
    return rec


def scene_to_patches(scene: np.ndarray, patch: int = 16, max_shift: int = 6) -> np.ndarray:
    """dataGenerator.py:108-121: [T,H,W] LR scene -> reflect pad max_shift/2 -> windows (patch+max_shift), stride patch.

    Returns [n*n, S, S, T, 1] in the model's channels-last layout (test.py:37-38 transpose).
    """
    T, H, W = scene.shape
    pad = max_shift // 2
    x = np.pad(scene, ((0, 0), (pad, pad), (pad, pad)), mode="reflect")
    S = patch + max_shift
    n = H // patch
    out = np.empty((n * n, S, S, T, 1), dtype=scene.dtype)
    k = 0
    for i in range(n):
        for j in range(n):
            w = x[:, i * patch:i * patch + S, j * patch:j * patch + S]        # [T,S,S]
            out[k, :, :, :, 0] = np.transpose(w, (1, 2, 0))
            k += 1
    return out
