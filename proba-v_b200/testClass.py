"""Predict path with the reference's surface: Enhancer(model, patchLR).enhance() (reference models/testClass.py:11-39)
and the functions test.py really uses -- evaluate / resolve / resolveByBatch / reconstruct_from_patches
(reference test.py:103-160).  Clip, round-half-even and the n x n stitch run on the device."""
from __future__ import annotations

import numpy as np


def resolve(model, lr_batch) -> np.ndarray:
    """test.py:114-122: model -> clip_by_value(0, 2**16) -> round -> numpy."""
    return model(lr_batch, resolve=True)


def resolveByBatch(model, lr_batch, batch_size=16) -> np.ndarray:
    """test.py:125-134 (patches are independent, so the batch split does not change the result)."""
    out = [resolve(model, lr_batch[s:s + batch_size]) for s in range(0, lr_batch.shape[0], batch_size)]
    return np.concatenate(out)


def reconstruct_from_patches(images: np.ndarray) -> np.ndarray:
    """test.py:149-160 host version (the device path is model.predict_scenes)."""
    n = int(len(images) ** 0.5)
    P = images.shape[1]
    rec = np.zeros((n * P, n * P, 1))
    k = 0
    for i in range(n):
        for j in range(n):
            rec[i * P:(i + 1) * P, j * P:(j + 1) * P] = images[k]
            k += 1
    return rec


def evaluate(model, X_test_patches):
    """test.py:103-111: list of stitched [n*P, n*P, 1] predictions, one per scene; one batched device call."""
    return list(model.predict_scenes(X_test_patches).astype(np.float64))


class Enhancer:
    def __init__(self, model, patchLR):
        self.model = model
        self.patchLR = patchLR

    def enhance(self):
        return evaluate(self.model, np.asarray(self.patchLR))

    def enhancePatch(self, set_):
        return resolve(self.model, set_)

    def reconstruct(self, patches):
        return reconstruct_from_patches(np.asarray(patches))


def calcRelativePSNR(patchPredOne, patchPredTwo, patchHR):
    """evaluate.py:76-87: shift-compensated cPSNR of two candidate reconstructions against the same (masked-array) HR scenes,
    [N, P, P, 1] each -> two [N] arrays.  The mask convention is the reference's: clear = ~patchHR.mask."""
    from .loss import Losses
    P = np.asarray(patchPredOne).shape[2]
    loss = Losses(targetShape=(P, P, 1))
    clear = ~np.ma.getmaskarray(patchHR)
    hr = np.asarray(np.ma.getdata(patchHR), np.float32)
    one = loss.shiftCompensatedcPSNR(hr, clear, np.asarray(patchPredOne, np.float32))
    two = loss.shiftCompensatedcPSNR(hr, clear, np.asarray(patchPredTwo, np.float32))
    return np.asarray(one), np.asarray(two)


def resolveBySampleAveraging(model, lr_batch, repeats: int = 20):
    """test.py:137-146: test-time augmentation over the frame order -- `repeats` cumulative random permutations of the T axis
    (np.random.permutation, the reference's global RNG), each resolved (clip + round), then averaged.  [B,S,S,T,1] -> [B,sP,sP,1]."""
    lr_batch = np.asarray(lr_batch, np.float32)
    acc = None
    for _ in range(repeats):
        newIdx = np.random.permutation(lr_batch.shape[3])
        lr_batch = lr_batch[:, :, :, newIdx, :]
        res = np.asarray(resolve(model, np.ascontiguousarray(lr_batch)), np.float64)
        acc = res if acc is None else acc + res
    return (acc / repeats).astype(np.float32)
