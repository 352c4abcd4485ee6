"""A second, independent restatement of the reference graph in plain NumPy (explicit tap loops written from the TensorFlow
op definitions, no torch ops), used to cross-check oracle/wdsr.py end to end on a tiny configuration.

Reference semantics restated here (citations are to /root/reference):
  * Keras Conv3D / Conv2D, channels-last, stride 1, cross-correlation: out[p] = sum_tap x[p + tap - pad] . w[tap]    (modelsTF.py:191-197)
  * 'same' padding = zero pad of k // 2 per side for odd k; 'valid' = none
  * TFA WeightNormalization(data_init=False): kernel = g * v / ||v||, norm over every axis but the last            (modelsTF.py:192,196)
  * tf.pad(mode='reflect') (no edge repeat), tf.nn.depth_to_space (NHWC), tf.reduce_mean over axis 3 (T)           (modelsTF.py:23,52,73,157)
  * graph wiring of WDSRConv3D.build                                                                               (modelsTF.py:15-74,152-189)
"""
import numpy as np
import pytest
import torch

from oracle.wdsr import OracleWDSR, init_params


def np_wn(v, g):
    axes = tuple(range(v.ndim - 1))
    return v * (g / np.sqrt(np.maximum((v * v).sum(axis=axes), 1e-12)))


def np_conv(x, w, b, same, relu):
    """x [B, *sp, Cin], w [*k, Cin, Cout]; loops over the kernel taps, vectorised over space."""
    nd = x.ndim - 2
    k = w.shape[:nd]
    if same:
        x = np.pad(x, [(0, 0)] + [(kk // 2, kk // 2) for kk in k] + [(0, 0)])
    out_sp = tuple(x.shape[1 + a] - (k[a] - 1) for a in range(nd))
    y = np.zeros((x.shape[0], *out_sp, w.shape[-1]))
    for tap in np.ndindex(*k):
        sl = (slice(None),) + tuple(slice(tap[a], tap[a] + out_sp[a]) for a in range(nd)) + (slice(None),)
        y += x[sl] @ w[tap]                    # [.., Cin] @ [Cin, Cout]
    y = y + b
    return np.maximum(y, 0.0) if relu else y


def np_depth_to_space(x, bs):
    B, H, W, C = x.shape
    co = C // (bs * bs)
    out = np.zeros((B, H * bs, W * bs, co))
    for i in range(bs):
        for j in range(bs):
            out[:, i::bs, j::bs, :] = x[:, :, :, (i * bs + j) * co:(i * bs + j + 1) * co]
    return out


def np_forward(p, x, mean, std, scale, R, T, P):
    L = lambda n, h, same, relu: np_conv(h, np_wn(p[n + "/v"], p[n + "/g"]), p[n + "/bias"], same, relu)   # noqa: E731
    mn = (x.mean(axis=3) - mean) / std
    h = L("mainConv1", (x - mean) / std, True, True)
    for i in range(R):
        h = L(f"normConv_{i}", L(f"decConv_{i}", L(f"expConv_{i}", h, True, True), True, False), True, False) + h
    nred, padded = {7: (2, ()), 9: (3, (1,)), 13: (5, (1, 2, 3))}[T]
    for i in range(1, nred + 1):
        if i in padded:
            h = np.pad(h, [(0, 0), (1, 1), (1, 1), (0, 0), (0, 0)], mode="reflect")
        h = L(f"convReducer_{i}", h, False, True)
    main = np_depth_to_space(L("upscaleConv1", h, False, False).reshape(x.shape[0], P, P, scale * scale), scale)
    r = mn
    for i in range(scale):
        r = L(f"residConv{i + 1}", r, False, i == 0)
    return (main + np_depth_to_space(r, scale)) * std + mean


@pytest.mark.parametrize("T", [9, 7, 13])
def test_numpy_twin_agrees_with_the_torch_oracle(T):
    scale, F, R, exp, dec, P, shift = 3, 4, 2, 2, 0.8, 4, 6
    om = OracleWDSR(8075.2045, 3160.7272, shift, scale, F, (3, 3, 3), R, exp, dec, T, P, True)
    p = init_params(om.specs, seed=T)
    rng = np.random.default_rng(T)
    x = rng.uniform(4000, 12000, size=(2, P + shift, P + shift, T, 1))
    ref = om.forward(p, torch.from_numpy(x)).numpy()
    got = np_forward({k: v.numpy() for k, v in p.items()}, x, om.mean, om.std, scale, R, T, P)
    assert got.shape == ref.shape == (2, scale * P, scale * P, 1)
    assert np.abs(got - ref).max() < 1e-8 * np.abs(ref).max()


def test_numpy_twin_depth_to_space_and_reflect_known_answers():
    x = np.arange(2 * 2 * 9, dtype=np.float64).reshape(1, 2, 2, 9)
    y = np_depth_to_space(x, 3)
    assert y.shape == (1, 6, 6, 1)
    for h in range(2):
        for w in range(2):
            for i in range(3):
                for j in range(3):
                    assert y[0, 3 * h + i, 3 * w + j, 0] == x[0, h, w, 3 * i + j]
    assert np.pad(np.array([1, 2, 3, 4]), 1, mode="reflect").tolist() == [2, 1, 2, 3, 4, 3]   # tf.pad REFLECT: no edge repeat


# ------------------------------------------------------------------------------------------------- losses + one backward
# Independent NumPy restatement of the shift-compensated L1 loss / cPSNR metric (reference models/loss.py:37-53,73-84,140-152,
# 168-187,226-238; utils/utils.py:42-44) with explicit Python loops over samples and shifts, and of its gradient by CENTRAL
# FINITE DIFFERENCES (no autograd, no closed form) -- cross-checks oracle/losses.py's values and its backward.
def np_shift_scores(hr, mask, sr, border=3):
    B, H, W, _ = hr.shape
    ch, cw = H - 2 * border, W - 2 * border
    S = 2 * border + 1
    l1 = np.zeros((S * S, B))
    cps = np.zeros((S * S, B))
    for b in range(B):
        pred = sr[b, border:border + ch, border:border + cw, 0].astype(np.float64)           # cropPrediction, loss.py:75-76
        for i in range(S):
            for j in range(S):
                h = hr[b, i:i + ch, j:j + cw, 0].astype(np.float64)                         # loss.py:141
                m = mask[b, i:i + ch, j:j + cw, 0].astype(np.float64)                       # loss.py:142
                n = m.sum()                                                                  # loss.py:144
                bias = (h - pred * m).sum() / n                                              # loss.py:182-187 (HR un-masked)
                corr = (pred + bias) * m                                                     # loss.py:148-149
                l1[i * S + j, b] = np.abs(h - corr).sum() / n                                # loss.py:226-228
                l2 = ((h - corr) ** 2).sum() / n
                cps[i * S + j, b] = 10.0 * np.log10(65535.0 ** 2 / l2)                       # loss.py:234-238
    return l1, cps


def np_shift_l1_loss(hr, mask, sr):
    return np_shift_scores(hr, mask, sr)[0].min(axis=0).mean()                              # loss.py:83-84


def test_numpy_loss_twin_agrees_with_the_torch_oracle_values_and_backward():
    from oracle.losses import OracleLosses
    rng = np.random.default_rng(5)
    B = 3
    hr = np.round(rng.random((B, 48, 48, 1)) * 4000 + 6000)
    sr = np.roll(hr, (2, -1), (1, 2)) + rng.standard_normal((B, 48, 48, 1)) * 45
    mask = rng.random((B, 48, 48, 1)) > 0.12                                               # raw HR stays under unclear pixels
    L = OracleLosses((48, 48, 1))
    t = lambda a: torch.from_numpy(np.asarray(a, np.float64))                               # noqa: E731
    l1, cps = np_shift_scores(hr, mask, sr)
    stack = L.stack("l1", t(hr), torch.from_numpy(mask), t(sr))[0].numpy()
    assert np.abs(stack - l1).max() < 1e-9 * np.abs(l1).max()
    assert np.array_equal(stack.argmin(axis=0), l1.argmin(axis=0))
    assert np.abs(L.shiftCompensatedcPSNR(t(hr), torch.from_numpy(mask), t(sr)).numpy() - cps.max(axis=0)).max() < 1e-9
    assert abs(float(L.shiftCompensatedL1Loss(t(hr), torch.from_numpy(mask), t(sr))) - np_shift_l1_loss(hr, mask, sr)) < 1e-9
    # backward: autograd of the oracle vs central differences of the NumPy twin on 40 random SR pixels (inside the crop and outside)
    srg = t(sr).clone().requires_grad_(True)
    L.shiftCompensatedL1Loss(t(hr), torch.from_numpy(mask), srg).backward()
    g = srg.grad.numpy()
    eps = 1e-3                                            # far below the distance of any residual to its kink on this data
    for _ in range(40):
        b, y, x = int(rng.integers(B)), int(rng.integers(48)), int(rng.integers(48))
        d = np.zeros_like(sr)
        d[b, y, x, 0] = eps
        fd = (np_shift_l1_loss(hr, mask, sr + d) - np_shift_l1_loss(hr, mask, sr - d)) / (2 * eps)
        assert abs(fd - g[b, y, x, 0]) < 1e-6 + 1e-4 * abs(g[b, y, x, 0]), (b, y, x, fd, g[b, y, x, 0])
    assert np.abs(g[:, :3]).max() == 0 and np.abs(g[:, :, 45:]).max() == 0                  # the 3-pixel border never sees a gradient
