"""Seeded synthetic PROBA-V-shaped data (no dataset / network in this environment).

LR: fp32, uint16 range, T registered (sub-pixel shifted) noisy 3x3-box down-samplings of one smooth scene.
HR: integer-valued fp32 patch of the clean scene, offset by a random integer shift in [-3, 3] so the
    shift search has a non-trivial answer.  mask: bool, True = clear (train.py:43).
"""
from __future__ import annotations

import numpy as np

BAND_STATS = {"NIR": (8075.2045, 3160.7272), "RED": (5266.2245, 3431.8614)}   # train.py:47-52


def _smooth_field(rng, n, size, cutoff=0.08):
    f = rng.standard_normal((n, size, size))
    F = np.fft.rfft2(f)
    fy = np.fft.fftfreq(size)[:, None]
    fx = np.fft.rfftfreq(size)[None, :]
    F *= np.exp(-(fy ** 2 + fx ** 2) / (2 * cutoff ** 2))
    g = np.fft.irfft2(F, s=(size, size))
    g /= g.std(axis=(1, 2), keepdims=True) + 1e-12
    return g


def make_masks(rng, B, H, W, clear_fraction=0.6, max_cover=0.15):
    m = np.ones((B, H, W, 1), dtype=bool)
    for b in range(B):
        if rng.random() < clear_fraction:
            continue
        budget = rng.uniform(0.01, max_cover) * H * W
        for _ in range(rng.integers(1, 4)):
            h = int(rng.integers(2, max(3, int(budget ** 0.5))))
            w = int(max(1, min(W, budget / (3 * h))))
            y = int(rng.integers(0, H - h + 1))
            x = int(rng.integers(0, W - w + 1))
            m[b, y:y + h, x:x + w, 0] = False
    return m


def make_batch(B, T=9, patch=16, scale=3, max_shift=6, band="NIR", seed=0, hr_zero_under_mask=False,
               all_clear=False):
    """Returns (lr [B,S,S,T,1] f32, hr [B,sP,sP,1] f32, mask [B,sP,sP,1] bool)."""
    rng = np.random.default_rng(seed)
    mean, std = BAND_STATS[band]
    S = patch + max_shift
    hs = S * scale                      # HR-equivalent side of the LR patch
    margin = 4
    size = hs + 2 * margin
    field = mean + 0.35 * std * _smooth_field(rng, B, size) + 0.03 * std * rng.standard_normal((B, size, size))
    field = np.clip(field, 0, 65535)
    lr = np.empty((B, S, S, T, 1), np.float32)
    for t in range(T):
        dy, dx = rng.integers(-1, 2, size=2)
        win = field[:, margin + dy:margin + dy + hs, margin + dx:margin + dx + hs]
        low = win.reshape(B, S, scale, S, scale).mean(axis=(2, 4))
        low = low + 0.01 * std * rng.standard_normal(low.shape)
        lr[:, :, :, t, 0] = np.clip(low, 0, 65535)
    P = patch * scale
    off = (hs - P) // 2
    sy, sx = rng.integers(-3, 4, size=(2, B))
    hr = np.empty((B, P, P, 1), np.float32)
    for b in range(B):
        y0, x0 = margin + off + sy[b], margin + off + sx[b]
        hr[b, :, :, 0] = np.round(field[b, y0:y0 + P, x0:x0 + P])
    mask = np.ones((B, P, P, 1), bool) if all_clear else make_masks(rng, B, P, P)
    if hr_zero_under_mask:
        hr = hr * mask
    return lr, hr.astype(np.float32), mask


def make_scene(ns, T=9, H=128, scale=3, band="NIR", seed=0):
    """Returns (lr scenes [ns,T,H,H] f32, hr scenes [ns,sH,sH,1] f32, mask [ns,sH,sH,1] bool)."""
    rng = np.random.default_rng(seed)
    mean, std = BAND_STATS[band]
    hs = H * scale
    field = mean + 0.35 * std * _smooth_field(rng, ns, hs + 8, cutoff=0.03) + 0.03 * std * rng.standard_normal((ns, hs + 8, hs + 8))
    field = np.clip(field, 0, 65535)
    lr = np.empty((ns, T, H, H), np.float32)
    for t in range(T):
        dy, dx = rng.integers(-1, 2, size=2)
        win = field[:, 4 + dy:4 + dy + hs, 4 + dx:4 + dx + hs]
        low = win.reshape(ns, H, scale, H, scale).mean(axis=(2, 4)) + 0.01 * std * rng.standard_normal((ns, H, H))
        lr[:, t] = np.clip(low, 0, 65535)
    hr = np.round(field[:, 4:4 + hs, 4:4 + hs])[..., None].astype(np.float32)
    mask = make_masks(rng, ns, hs, hs, clear_fraction=0.3, max_cover=0.1)
    return lr, hr, mask
