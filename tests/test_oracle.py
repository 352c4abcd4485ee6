"""CPU: pins the oracle with hand-derived known answers and torch cross-checks (the reference ships no tests,
golden vectors or weights, and TensorFlow is not installed: PARITY UNPINNED -- see oracle/__init__.py)."""
import math

import numpy as np
import pytest
import torch

from oracle import cshift
from oracle.losses import OracleLosses, sobel_edges
from oracle.optim import OracleAdam, OracleNadam
from oracle.step import reconstruct_from_patches, scene_to_patches
from oracle.wdsr import OracleWDSR, conv_cl, depth_to_space, init_params, reflect_pad_hwt, wn_kernel
from tests.helpers import NIR


def test_variable_inventory_matches_checkpoint_index():
    # SURVEY Appendix D (modelInfo/ckpt_p16t9c85r12/NIR/ckpt-124.index): 44 layers, 535 267 parameters
    m = OracleWDSR(*NIR, 6, 3, 32, (3, 3, 3), 12, 8, 0.8, 9, 16)
    assert len(m.specs) == 44
    assert m.n_params() == 535267
    shapes = {s["name"]: (*s["k"], s["cin"], s["cout"]) for s in m.specs}
    assert shapes["mainConv1"] == (3, 3, 3, 1, 32)
    assert shapes["expConv_0"] == (1, 1, 1, 32, 256)
    assert shapes["decConv_11"] == (1, 1, 1, 256, 25)
    assert shapes["normConv_5"] == (3, 3, 3, 25, 32)
    assert shapes["convReducer_3"] == (3, 3, 3, 32, 32)
    assert shapes["upscaleConv1"] == (3, 3, 3, 32, 9)
    assert shapes["residConv1"] == (3, 3, 1, 9) and shapes["residConv3"] == (3, 3, 9, 9)


def test_macs_match_survey_appendix_a():
    assert OracleWDSR(*NIR, 6, 3, 32, (3, 3, 3), 12, 8, 0.8, 9, 16).macs_per_patch() == 2073878964
    assert OracleWDSR(*NIR, 6, 3, 32, (3, 3, 3), 12, 8, 0.8, 13, 16).macs_per_patch() == 3183996852
    m = OracleWDSR(*NIR, 6, 3, 64, (3, 3, 3), 24, 8, 0.8, 9, 16)
    assert m.macs_per_patch() == 16084133172 and m.n_params() == 3909467


def test_shape_contract_readme():
    # README.md:223-224: 22x22 LR patches <-> 48x48 HR patches
    for T in (7, 9, 13, 19):
        m = OracleWDSR(*NIR, 6, 3, 8, (3, 3, 3), 1, 2, 0.8, T, 16)
        p = init_params(m.specs, 0)
        y = m.forward(p, torch.rand(2, 22, 22, T, 1, dtype=torch.float64) * 1e4)
        assert y.shape == (2, 48, 48, 1)
    with pytest.raises(ValueError):
        OracleWDSR(*NIR, 6, 3, 8, (3, 3, 3), 1, 2, 0.8, 12, 16)


def test_weight_norm_identity_and_torch_equivalent():
    torch.manual_seed(0)
    v = torch.randn(3, 3, 3, 5, 7, dtype=torch.float64)
    nrm = torch.sqrt((v * v).reshape(-1, 7).sum(0))
    assert torch.allclose(wn_kernel(v, nrm), v, atol=1e-14)          # g = ||v||  =>  w = v  (Appendix B.1)
    g = torch.rand(7, dtype=torch.float64) + 0.5
    conv = torch.nn.utils.weight_norm(torch.nn.Conv3d(5, 7, 3, bias=False).double(), dim=0)
    conv.weight_v.data = v.permute(4, 3, 0, 1, 2).contiguous()
    conv.weight_g.data = g.reshape(7, 1, 1, 1, 1)
    x = torch.randn(2, 6, 6, 6, 5, dtype=torch.float64)
    ref = conv(x.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
    got = conv_cl(x, wn_kernel(v, g), torch.zeros(7, dtype=torch.float64), "valid", False)
    assert torch.allclose(got, ref, atol=1e-12)


def test_depth_to_space_is_pixel_shuffle_for_one_channel():
    x = torch.arange(2 * 4 * 4 * 9, dtype=torch.float64).reshape(2, 4, 4, 9)
    ref = torch.nn.functional.pixel_shuffle(x.permute(0, 3, 1, 2), 3).permute(0, 2, 3, 1)
    assert torch.equal(depth_to_space(x, 3), ref)
    assert depth_to_space(x, 3)[1, 3 * 2 + 1, 3 * 3 + 2, 0] == x[1, 2, 3, 1 * 3 + 2]      # Appendix B.3


def test_reflect_pad_matches_numpy():
    x = torch.rand(1, 5, 6, 4, 2, dtype=torch.float64)
    ref = np.pad(x.numpy(), ((0, 0), (2, 2), (1, 1), (1, 1), (0, 0)), mode="reflect")
    assert np.array_equal(reflect_pad_hwt(x, 2, 1, 1).numpy(), ref)


def test_sobel_known_answer():
    img = torch.zeros(1, 5, 5, 1, dtype=torch.float64)
    img[0, :, 3:, 0] = 1.0                                           # vertical step edge
    e = sobel_edges(img)
    assert e.shape == (1, 5, 5, 1, 2)
    assert torch.all(e[..., 0] == 0)                                  # dy = 0 everywhere
    assert e[0, 2, 2, 0, 1] == 4 and e[0, 2, 3, 0, 1] == 4 and e[0, 2, 0, 0, 1] == 0


def _rand_loss_inputs(B=6, seed=0, all_clear=False):
    g = torch.Generator().manual_seed(seed)
    hr = torch.round(torch.rand(B, 48, 48, 1, generator=g, dtype=torch.float64) * 4000 + 6000)
    sr = (hr + torch.randn(B, 48, 48, 1, generator=g, dtype=torch.float64) * 50).float().double()   # fp32-representable
    mask = torch.ones(B, 48, 48, 1, dtype=torch.bool) if all_clear else torch.rand(B, 48, 48, 1, generator=g) > 0.1
    return hr, mask, sr


def test_loss_known_answer_shift_and_bias():
    # SR crop == HR window (5,1) + const, all clear  =>  loss 0 at stack index 5*7+1 = 36  
    g = torch.Generator().manual_seed(3)
    hr = torch.round(torch.rand(1, 48, 48, 1, generator=g, dtype=torch.float64) * 4000 + 6000)
    sr = torch.zeros_like(hr)
    sr[:, 3:45, 3:45] = hr[:, 5:47, 1:43] + 123.0
    mask = torch.ones(1, 48, 48, 1, dtype=torch.bool)
    L = OracleLosses((48, 48, 1))
    best, idx, cnt, stack = L.details("l1", hr, mask, sr)
    assert int(idx[0]) == 36 and float(best[0]) < 1e-9 and int(cnt[0]) == 42 * 42
    assert float(L.shiftCompensatedL1Loss(hr, mask, sr)) < 1e-9
    _, _, b = L.stack("l1", hr, mask, sr)
    assert abs(float(b[36, 0]) + 123.0) < 1e-9


def test_cpsnr_known_mse():
    hr = torch.full((1, 48, 48, 1), 1000.0, dtype=torch.float64)
    sr = hr.clone()
    sr[:, 3:45:2, 3:45, :] += 20.0      # half the rows +20: after bias correction residual is +-10 => MSE 100
    mask = torch.ones(1, 48, 48, 1, dtype=torch.bool)
    c = OracleLosses((48, 48, 1)).shiftCompensatedcPSNR(hr, mask, sr)
    assert abs(float(c[0]) - 10 * math.log10(65535.0 ** 2 / 100.0)) < 1e-9


def test_hr_is_not_masked_quirk():
    # loss.py:140-152: only SR is multiplied by the mask; unclear HR pixels enter sum(h) and |r|
    hr, mask, sr = _rand_loss_inputs(2, seed=5)
    L = OracleLosses((48, 48, 1))
    a = L.shiftCompensatedL1Loss(hr, mask, sr)
    b = L.shiftCompensatedL1Loss(hr * mask, mask, sr)
    assert abs(float(a) - float(b)) > 1.0


def test_c_oracle_agrees_with_torch_oracle():
    hr, mask, sr = _rand_loss_inputs(4, seed=7)
    L = OracleLosses((48, 48, 1))
    for kind, name in ((0, "l1"), (1, "l2"), (2, "cpsnr"), (3, "l1edge")):
        s, n, b = L.stack(name, hr, mask, sr)
        sc, cn, bi = cshift.shift_scores(kind, hr[..., 0].numpy().astype(np.float32), mask[..., 0].numpy(),
                                         sr[..., 0].numpy().astype(np.float32))
        assert np.allclose(sc, s.T.numpy(), rtol=1e-9, atol=1e-9), name
        assert np.array_equal(cn, n.T.numpy())
        assert np.allclose(bi, b.T.numpy(), rtol=1e-9, atol=1e-9)


def test_l1_closed_form_gradient_matches_autograd():
    hr, mask, sr = _rand_loss_inputs(3, seed=11)
    L = OracleLosses((48, 48, 1))
    srg = sr.clone().requires_grad_(True)
    L.shiftCompensatedL1Loss(hr, mask, srg).backward()
    cf = L.l1_grad_closed_form(hr, mask, sr)
    assert torch.allclose(srg.grad, cf, atol=1e-15)
    assert torch.all(srg.grad[:, :3] == 0) and torch.all(srg.grad[:, :, 45:] == 0)      # 3-px border gets no gradient


def test_nadam_matches_torch_nadam():
    torch.manual_seed(0)
    w0 = torch.randn(50, dtype=torch.float64)
    p_t = w0.clone().requires_grad_(True)
    opt = torch.optim.NAdam([p_t], lr=5e-4, betas=(0.9, 0.999), eps=1e-7, momentum_decay=4e-3)
    mine = OracleNadam(5e-4)
    params = {"w": w0.clone()}
    for k in range(5):
        g = torch.randn(50, dtype=torch.float64, generator=torch.Generator().manual_seed(k))
        p_t.grad = g.clone()
        opt.step()
        params = mine.apply_gradients(params, {"w": g})
    assert torch.allclose(params["w"], p_t.detach(), atol=1e-12)


def test_adam_matches_torch_adam():
    w0 = torch.linspace(-1, 1, 20, dtype=torch.float64)
    p_t = w0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_t], lr=1e-3, eps=1e-7)
    mine = OracleAdam(1e-3)
    params = {"w": w0.clone()}
    for k in range(4):
        g = torch.randn(20, dtype=torch.float64, generator=torch.Generator().manual_seed(k))
        p_t.grad = g.clone()
        opt.step()
        params = mine.apply_gradients(params, {"w": g})
    # Keras folds the bias correction into lr_t and adds eps outside the corrected sqrt: tiny, bounded difference
    assert torch.allclose(params["w"], p_t.detach(), atol=1e-8)


def test_scene_patch_geometry_roundtrip():
    # dataGenerator.py:108-121 + test.py:149-160: the 16x16 cores of the 64 patches tile the scene exactly
    rng = np.random.default_rng(0)
    scene = rng.random((9, 128, 128)).astype(np.float32)
    patches = scene_to_patches(scene)
    assert patches.shape == (64, 22, 22, 9, 1)
    cores = patches[:, 3:19, 3:19, 4, :]
    assert np.array_equal(reconstruct_from_patches(cores)[..., 0], scene[4])
    assert patches[0, 0, 0, 0, 0] == scene[0, 3, 3]                  # reflect pad: index -3 -> 3
