"""CPU pins for the arithmetic claims behind the pair-row formats of the error-compensated engine (csrc/rows.h, DESIGN.md sections 2a / 3):
what the packed fp16 pair rows (forward) and the bf16 pair rows (backward) can represent, and how close the three-product forward form
gets to the exact product.  NumPy emulation of the operand roundings; products are exact and sums run in fp64 (the tensor core's own
fp32 accumulation is measured on the GPU by the self-test)."""
import numpy as np

PACK_SCALE = 4096.0


def tf32_rn(x):
    """round-to-nearest (ties away, like cvt.rna.tf32.f32) to 10 explicit mantissa bits"""
    u = np.asarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16_rn(x):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def fp16_pair(v):
    """rows.h: [fp16(hi) | fp16(2^12 lo')] with hi = tf32(v), lo' = v - fp16(hi)"""
    v = np.asarray(v, np.float32)
    h16 = tf32_rn(v).astype(np.float16)
    lo = (v - h16.astype(np.float32)).astype(np.float32)
    l16 = (lo * np.float32(PACK_SCALE)).astype(np.float16)
    return h16, l16


def test_fp16_pair_row_reconstructs_the_fp32_value():
    rng = np.random.default_rng(0)
    # O(1) activations, values down to the edge of fp16's normal range, and below it (subnormal hi halves)
    v = np.concatenate([rng.standard_normal(200_000).astype(np.float32) * s for s in (1.0, 30.0, 1e-2, 1e-4, 3e-6)])
    h16, l16 = fp16_pair(v)
    recon = h16.astype(np.float64) + l16.astype(np.float64) / PACK_SCALE
    err = np.abs(recon - v.astype(np.float64))
    # 11 bits in each half: 2^-23 of |v| (one more factor 2 where the lo half itself is subnormal: |lo| 2^12 < 6.1e-5)
    assert np.all(err <= 2.0 ** -22 * np.abs(v) + 2.0 ** -24 / PACK_SCALE)
    big = np.abs(v) > 1e-3
    assert np.max(err[big] / np.abs(v[big])) <= 2.0 ** -22
    # in fp16's normal range the hi half IS the tf32 value (the single-pass engine's operand)
    normal = np.abs(v) > 6.2e-5
    assert np.array_equal(h16[normal].astype(np.float32), tf32_rn(v)[normal])


def test_bf16_pair_row_keeps_16_bits_and_the_exponent_range():
    rng = np.random.default_rng(1)
    g = (rng.standard_normal(500_000) * np.exp(rng.uniform(-60, 10, 500_000))).astype(np.float32)     # gradients: any magnitude
    a = bf16_rn(g)
    b = bf16_rn(g - a)
    err = np.abs((a.astype(np.float64) + b.astype(np.float64)) - g.astype(np.float64))
    ok = np.abs(g) > 1e-30                                                                          # (fp32 subnormals aside)
    assert np.max(err[ok] / np.abs(g[ok])) <= 2.0 ** -16
    # fp16 could not hold them: most of these values are below its smallest normal number
    assert np.mean(np.abs(g) < 6.1e-5) > 0.5


def test_three_product_form_is_fp32_grade_and_single_pass_is_not():
    """x w ~= x_hi w_hi + x_lo w_hi + x_hi w_lo from the packed rows, K = 27 * 32 (one output of a 3x3x3 conv)."""
    rng = np.random.default_rng(2)
    K, n = 864, 4000
    x = rng.standard_normal((n, K)).astype(np.float32)
    w = (rng.standard_normal((n, K)) * 0.05).astype(np.float32)
    exact = np.sum(x.astype(np.float64) * w.astype(np.float64), axis=1)
    xh, xl = fp16_pair(x)
    wh, wl = fp16_pair(w)
    xh, xl, wh, wl = (t.astype(np.float64) for t in (xh, xl, wh, wl))
    main = np.sum(xh * wh, axis=1)
    corr = np.sum(xl * wh + xh * wl, axis=1) / PACK_SCALE
    scale = np.sqrt(np.sum((x.astype(np.float64) * w.astype(np.float64)) ** 2, axis=1))            # rms size of the summands
    e3 = np.abs(main + corr - exact) / scale
    e1 = np.abs(main - exact) / scale
    # what is dropped is x_lo w_lo (2^-24 per term) and the 11-bit rounding of the lo halves (measured: median 5e-8, max 3e-7)
    assert np.max(e3) < 5e-7 and np.median(e3) < 1e-7
    # single pass (the hi halves alone = the tf32 product): the error the L1 signs react to (measured: median 2e-4)
    assert np.median(e1) > 1e-4
    assert np.median(e1) > 1000 * np.median(e3)


def test_bf16_pair_data_gradient_drops_2_to_the_minus_16():
    """g (w_hi + w_lo) as g_a w_a + g_b w_a + g_a w_b with bf16 pairs (conv3_tc.cu MODE 2): per-term error ~2^-16, no bias."""
    rng = np.random.default_rng(3)
    K, n = 864, 4000
    g = (rng.standard_normal((n, K)) * 1e-6).astype(np.float32)
    w = (rng.standard_normal((n, K)) * 0.05).astype(np.float32)
    ga, wa = bf16_rn(g), bf16_rn(w)
    gb, wb = bf16_rn(g - ga), bf16_rn(w - wa)
    ga, gb, wa, wb = (t.astype(np.float64) for t in (ga, gb, wa, wb))
    got = np.sum(ga * wa + gb * wa + ga * wb, axis=1)
    exact = np.sum(g.astype(np.float64) * w.astype(np.float64), axis=1)
    scale = np.sqrt(np.sum((g.astype(np.float64) * w.astype(np.float64)) ** 2, axis=1))
    e = (got - exact) / scale
    assert np.max(np.abs(e)) < 2e-5 and abs(np.mean(e)) < 2e-7     # far below tf32's 2^-12 per operand, and unbiased
