"""GPU parity, edge cases of SURVEY.md section 4(3) / VERDICT round 1: windows without a clear pixel (N = 0), exact ties in
the min over shifts, property-based clearance masks, raw HR under the mask for every loss, the T = 19 graph, an empty
data-parallel shard.  Through the C-ABI, against the oracle (PARITY UNPINNED w.r.t. TensorFlow, see oracle/__init__.py).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle.losses import OracleLosses
from oracle.step import loss_and_grads
from tests.helpers import cuda_model, oracle_and_params, rel_err

pytestmark = pytest.mark.gpu


def _np(t, dt=np.float32):
    return t.detach().cpu().numpy().astype(dt)


def _inputs(B, seed, noise=50.0):
    g = torch.Generator().manual_seed(seed)
    hr = torch.round(torch.rand(B, 48, 48, 1, generator=g, dtype=torch.float64) * 4000 + 6000)
    sr = (hr.roll((1, -2), (1, 2)) + torch.randn(B, 48, 48, 1, generator=g, dtype=torch.float64) * noise).float().double()
    return hr, sr


# ------------------------------------------------------------------------------------------------- N = 0
def test_all_unclear_sample_gives_nan_like_the_reference():
    """loss.py:144-146,184: N = sum(mask window) is unguarded, so a sample without a clear pixel has b = (1/0) * ... and
    every score is NaN / inf; reduce_min then yields NaN for that sample and the batch mean is NaN.  Same here."""
    import probav_b200 as pb
    hr, sr = _inputs(3, 1)
    mask = torch.ones(3, 48, 48, 1, dtype=torch.bool)
    mask[1] = False
    L = OracleLosses((48, 48, 1))
    ref = L.stack("l1", hr, mask, sr)[0].min(dim=0).values
    assert torch.isnan(ref[1]) and torch.isfinite(ref[[0, 2]]).all()
    out = pb.Losses((48, 48, 1)).evaluate("l1", _np(hr), _np(mask, np.uint8), _np(sr), want_grad=True)
    assert np.isnan(out["loss_per_sample"][1]) and np.isnan(out["mean_loss"][0])
    assert rel_err(out["loss_per_sample"][[0, 2]], _np(ref[[0, 2]], np.float64)) < 1e-3
    assert out["clear_count"][1] == 0
    # the clear samples' gradients are unaffected by their NaN neighbour
    assert np.isfinite(out["dsr"][[0, 2]]).all()


def test_some_windows_without_clear_pixels_documented_divergence():
    """Only a corner of the sample is clear, so some of the 49 windows contain no clear pixel (NaN score) and others do.
    Reference / oracle: the NaN propagates through reduce_min (loss.py:83).  The CUDA kernel's arg-min skips NaN scores and
    returns the best FINITE shift (documented in DESIGN.md section 2 -- a sample the reference cannot train on, made usable;
    PROBA-V clearance masks never produce it: the data generator drops patches under 85 % clearance, cfg ckpt_dir p16t9c85*)."""
    import probav_b200 as pb
    hr, sr = _inputs(2, 2)
    mask = torch.zeros(2, 48, 48, 1, dtype=torch.bool)
    mask[:, 43:, 43:] = True                      # windows (i, j) reach rows/cols up to i + 41: clear pixels only when i, j >= 2
    L = OracleLosses((48, 48, 1))
    stack, cnt, _ = L.stack("l1", hr, mask, sr)
    assert (cnt == 0).any() and (cnt > 0).any()
    assert torch.isnan(stack.min(dim=0).values).all()          # the reference's answer
    out = pb.Losses((48, 48, 1)).evaluate("l1", _np(hr), _np(mask, np.uint8), _np(sr), want_stack=True)
    finite = torch.where(torch.isnan(stack), torch.full_like(stack, float("inf")), stack)
    idx = finite.argmin(dim=0)
    assert np.array_equal(out["best_shift"], _np(idx, np.int32))
    assert rel_err(out["loss_per_sample"], _np(finite.min(dim=0).values, np.float64)) < 1e-3
    assert np.array_equal(out["stack"][:, :, 2].astype(np.int64), _np(cnt.T, np.int64))


# ------------------------------------------------------------------------------------------------- ties
def test_exact_ties_take_the_first_minimum():
    """Constant HR and SR with an all-clear mask: all 49 shifts score exactly the same.  Policy: value and
    gradient at the FIRST minimum (shift index 0), as torch.min / argmin; TensorFlow's reduce_min gradient would split the
    gradient evenly over the tied shifts (loss.py:83) -- exact ties do not occur on real data."""
    import probav_b200 as pb
    hr = torch.full((2, 48, 48, 1), 7000.0, dtype=torch.float64)
    sr = torch.full((2, 48, 48, 1), 7100.0, dtype=torch.float64)
    sr[:, 10, 10] += 64.0                          # one off pixel so that the gradient is not identically zero
    mask = torch.ones(2, 48, 48, 1, dtype=torch.bool)
    L = OracleLosses((48, 48, 1))
    stack = L.stack("l1", hr, mask, sr)[0]
    assert float(stack.max() - stack.min()) == 0.0
    out = pb.Losses((48, 48, 1)).evaluate("l1", _np(hr), _np(mask, np.uint8), _np(sr), want_grad=True)
    assert np.array_equal(out["best_shift"], np.zeros(2, np.int32))
    ref = L.l1_grad_closed_form(hr, mask, sr)                   # evaluated at the first arg-min
    assert np.abs(out["dsr"] - _np(ref, np.float64)).max() <= 1e-3 * np.abs(_np(ref, np.float64)).max()


# ------------------------------------------------------------------------------------------------- property-based masks
def test_property_based_masks_match_oracle():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    import probav_b200 as pb
    L = OracleLosses((48, 48, 1))
    PL = pb.Losses((48, 48, 1))

    @st.composite
    def masks(draw):
        m = np.ones((2, 48, 48, 1), bool)
        for b in range(2):
            kind = draw(st.sampled_from(["rows", "cols", "blob", "one_clear_per_window", "speckle", "stripe"]))
            if kind == "rows":                     # whole rows unclear
                for r in draw(st.lists(st.integers(0, 47), min_size=1, max_size=20, unique=True)):
                    m[b, r] = False
            elif kind == "cols":
                for c in draw(st.lists(st.integers(0, 47), min_size=1, max_size=20, unique=True)):
                    m[b, :, c] = False
            elif kind == "blob":
                y, x, h, w = draw(st.integers(0, 40)), draw(st.integers(0, 40)), draw(st.integers(1, 30)), draw(st.integers(1, 30))
                m[b, y:y + h, x:x + w] = False
            elif kind == "one_clear_per_window":   # a single clear pixel that every 42x42 window contains
                m[b] = False
                m[b, draw(st.integers(6, 41)), draw(st.integers(6, 41))] = True
            elif kind == "speckle":
                rng = np.random.default_rng(draw(st.integers(0, 2 ** 16)))
                m[b] = rng.random((48, 48, 1)) > draw(st.floats(0.05, 0.9))
                m[b, 20, 20] = True
            else:
                m[b, :, ::draw(st.integers(2, 5))] = False
        return m

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(masks(), st.integers(0, 1000), st.sampled_from(["l1", "l2"]))
    def check(m, seed, kind):
        hr, sr = _inputs(2, seed)
        mask = torch.from_numpy(m)
        best, idx, cnt, stack = L.details(kind, hr, mask, sr)
        out = PL.evaluate(kind, _np(hr), m.astype(np.uint8), _np(sr), want_grad=True)
        # the index (and the clear count at it) must be bit-exact unless the two best scores are closer than fp32 can tell apart
        s = np.sort(_np(stack, np.float64), axis=0)
        for b in range(2):
            if (s[1, b] - s[0, b]) > 1e-5 * abs(s[0, b]):
                assert out["best_shift"][b] == int(idx[b])
            if out["best_shift"][b] == int(idx[b]):
                assert out["clear_count"][b] == int(cnt[b])
        # (a single clear pixel makes every residual exactly 0 in fp64 and a rounding-sized number in fp32: absolute floor)
        refl = _np(best, np.float64)
        assert (np.abs(out["loss_per_sample"] - refl) <= 1e-3 * np.abs(refl) + 1e-2).all(), (out["loss_per_sample"], refl)
        refc = _np(L.shiftCompensatedcPSNR(hr, mask, sr), np.float64)
        ok = np.isfinite(refc) & (refc < 100.0)
        assert np.abs(out["cpsnr"][ok] - refc[ok]).max(initial=0.0) < 0.01
        srg = sr.clone().requires_grad_(True)
        (L.shiftCompensatedL1Loss if kind == "l1" else L.shiftCompensatedL2Loss)(hr, mask, srg).backward()
        ref = _np(srg.grad, np.float64)
        if all(out["best_shift"][b] == int(idx[b]) for b in range(2)) and (refl > 1.0).all():
            # |r| has a kink at 0: a residual below fp32 resolution may take the other sign (a handful of pixels at most), and
            # every flipped sign moves the bias-correction term sum(s m) / N of ALL pixels of its sample by 2 / N
            # (unit of the metric: the natural gradient magnitude 1 / (N B) -- when raw HR under a few unclear pixels pushes the
            # bias so far that every residual has the same sign, the true gradient is exactly 0 and fp32 leaves ~1e-11)
            err = np.abs(out["dsr"] - ref) / max(np.abs(ref).max(), 1.0 / (1764 * 2))
            slack = 1e-3 + 8 * 2.0 / max(1, int(out["clear_count"].min()))
            assert (err > slack).sum() <= 8, (kind, seed, [(int((err[b] > slack).sum()), int(m[b].sum()), int(m[b, 3:45, 3:45].sum()), float(err[b].max()),
                                                            int(out["best_shift"][b]), float(refl[b]), float(np.abs(ref[b]).max())) for b in range(2)], slack)

    check()


# ------------------------------------------------------------------------------------------------- raw HR under the mask
@pytest.mark.parametrize("kind", ["sobel_l1_mix"])
def test_l1edge_with_raw_hr_under_unclear_pixels(kind):
    """The reference feeds np.array(masked_array), i.e. the raw HR values under unclear pixels (dataGenerator), and they enter
    the bias, the L1 and the Sobel term un-masked (loss.py:141-152).  ADVICE round 1: this path had no L1Edge coverage."""
    import probav_b200 as pb
    g = torch.Generator().manual_seed(77)
    hr, sr = _inputs(8, 78)
    mask = torch.rand(8, 48, 48, 1, generator=g) > 0.15
    L = OracleLosses((48, 48, 1))
    best, idx, cnt, _ = L.details("l1edge", hr, mask, sr)
    out = pb.Losses((48, 48, 1)).evaluate(kind, _np(hr), _np(mask, np.uint8), _np(sr), want_grad=True)
    assert np.array_equal(out["best_shift"], _np(idx, np.int32))
    assert np.array_equal(out["clear_count"], _np(cnt, np.int32))
    assert rel_err(out["loss_per_sample"], _np(best, np.float64)) < 1e-3
    srg = sr.clone().requires_grad_(True)
    L.shiftCompensatedL1EdgeLoss(hr, mask, srg).backward()
    ref = _np(srg.grad, np.float64)
    err = np.abs(out["dsr"] - ref) / np.abs(ref).max()
    assert (err > 1e-3).sum() <= 16 and err.max() < 1.0, ((err > 1e-3).sum(), err.max())


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32"])
def test_forward_backward_with_raw_hr(small_cfg, precision):
    import probav_b200 as pb
    import tempfile
    from probav_b200 import synth
    om, p = oracle_and_params(small_cfg, seed=60)
    m = cuda_model(small_cfg, p, precision=precision)
    lr, hr, mask = synth.make_batch(6, seed=61, hr_zero_under_mask=False)
    ol = OracleLosses((48, 48, 1))
    loss, g, sr, cps = loss_and_grads(om, ol, p, torch.from_numpy(lr).double(), torch.from_numpy(hr).double(), torch.from_numpy(mask))
    L = pb.Losses((48, 48, 1))
    d = tempfile.mkdtemp(prefix="pv_")
    t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/ckpt", d + "/log")
    lossv, psnrv = t.forward_backward(lr, hr, mask)
    assert abs(lossv - float(loss)) < 1e-3 * float(loss)
    assert abs(psnrv - float(cps.mean())) < 0.01
    got = t.get_grads()
    worst = max(rel_err(got[k], v.numpy()) for k, v in g.items() if np.abs(v.numpy()).max() > 0)
    print(f"{precision}: raw-HR gradients, worst {worst:.2e}")
    # Small batches are ill-conditioned for this metric: on this very case the ORACLE evaluated in fp32 (what TensorFlow
    # itself computes) is 1.98e-3 away from its fp64 self -- one L1 sign flip.  test_full_batch_gradients_match_golden is the
    # 1e-3 check, at batch 128.
    assert worst < {"fp32": 4e-3, "tf32x3": 4e-3, "tf32": 3e-2}[precision]


# ------------------------------------------------------------------------------------------------- T = 19
def test_t19_graph_forward_and_gradients(small_cfg):
    """ConvReduceAndUpscaleEx (modelsTF.py:76-121): 5x5x5 first reducer, reflect pads along T -- dense fp32 engine."""
    import probav_b200 as pb
    import tempfile
    from probav_b200 import synth
    cfg = dict(small_cfg, numImgLR=19, numResBlocks=1)
    om, p = oracle_and_params(cfg, seed=70)
    m = cuda_model(cfg, p, precision="fp32")
    lr, hr, mask = synth.make_batch(3, T=19, seed=71, hr_zero_under_mask=False)
    ol = OracleLosses((48, 48, 1))
    loss, g, sr, cps = loss_and_grads(om, ol, p, torch.from_numpy(lr).double(), torch.from_numpy(hr).double(), torch.from_numpy(mask))
    got = m(lr)
    assert rel_err(got, sr.numpy()) < 1e-3 and np.abs(got - sr.numpy()).max() / 3160.7272 < 1e-3
    L = pb.Losses((48, 48, 1))
    d = tempfile.mkdtemp(prefix="pv_")
    t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/ckpt", d + "/log")
    lossv, psnrv = t.forward_backward(lr, hr, mask)
    assert abs(lossv - float(loss)) < 1e-3 * float(loss) and abs(psnrv - float(cps.mean())) < 0.01
    grads = t.get_grads()
    assert set(k.split("/")[0] for k in grads if k.startswith("convReducer")) == {f"convReducer_{i}" for i in range(1, 11)}
    worst = max(rel_err(grads[k], v.numpy()) for k, v in g.items() if np.abs(v.numpy()).max() > 0)
    print(f"T=19: worst gradient rel err {worst:.2e}")
    assert worst < 1e-3
    with pytest.raises(ValueError):
        cuda_model(cfg, p, precision="tf32")       # the row engine has no 5x5x5 reducer


# ------------------------------------------------------------------------------------------------- empty data-parallel shard
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_empty_shard_contributes_zero_gradients(small_cfg, precision):
    """a rank whose shard of the last partial global batch is empty must not raise (the other ranks would hang in
    the all-reduce): pv_train_forward_backward[_staged] with B = 0 zero-fills its gradient range and reports it."""
    import probav_b200 as pb
    import tempfile
    from probav_b200 import _buf, _lib, synth
    from probav_b200._lib import check
    om, p = oracle_and_params(small_cfg, seed=80)
    m = cuda_model(small_cfg, p, precision=precision)
    L = pb.Losses((48, 48, 1))
    d = tempfile.mkdtemp(prefix="pv_")
    t = pb.ModelTrainer(m, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/ckpt", d + "/log")
    lr, hr, mask = synth.make_batch(2, seed=81)
    t.forward_backward(lr, hr, mask)
    assert float(t.grad_view().abs().max()) > 0
    dev = torch.device(f"cuda:{m.device}")
    out = torch.full((2,), 7.0, dtype=torch.float32, device=dev)
    lo, hi = C.c_int64(), C.c_int64()
    covered = 0
    for stage in (0, 1):
        check(_lib.lib().pv_train_forward_backward_staged(t._h, None, None, None, 0, 1.0, _buf.ptr(out), stage, C.byref(lo), C.byref(hi),
                                                          _buf.current_stream_ptr(dev)))
        covered += hi.value - lo.value
    torch.cuda.synchronize()
    assert covered == m.nparams and float(t.grad_view().abs().max()) == 0.0 and float(out.abs().max()) == 0.0
    t.forward_backward(lr, hr, mask)
    check(_lib.lib().pv_train_forward_backward(t._h, None, None, None, 0, 1.0, _buf.ptr(out), _buf.current_stream_ptr(dev)))
    torch.cuda.synchronize()
    assert float(t.grad_view().abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------- PyTorch twin
def test_pytorch_twin_matches_the_trainer_and_steps_with_torch_optim(small_cfg):
    """proba-v_b200/modelsPyTorch.py (the working counterpart of the reference's draft models/modelsPyTorch.py:10-151): the nn.Module
    runs the same kernels, its autograd gradients equal ModelTrainer's tape.gradient, and a torch optimizer steps the engine's
    weight arena in place."""
    import tempfile
    import probav_b200 as pb
    from probav_b200 import synth
    from probav_b200.modelsPyTorch import Conv3DResNet, ShiftL1Loss
    om, p = oracle_and_params(small_cfg, seed=90)
    lr, hr, mask = synth.make_batch(4, seed=91, hr_zero_under_mask=False)
    net = Conv3DResNet((1, 9, 22, 22), 3, small_cfg["numResBlocks"], (3, 3, 3), 32, 8, 0.8, precision="tf32x3")
    net.model.set_weights({k: v.numpy().astype(np.float32) for k, v in p.items()})
    x = torch.from_numpy(lr).cuda().permute(0, 4, 3, 1, 2).contiguous()                 # [B, 1, T, H, W]
    y = torch.from_numpy(hr).cuda().permute(0, 3, 1, 2)
    k = torch.from_numpy(mask).cuda().permute(0, 3, 1, 2)
    crit = ShiftL1Loss("l1")
    sr = net(x)
    assert tuple(sr.shape) == (4, 1, 48, 48)
    ref_sr = om.forward(p, torch.from_numpy(lr).double()).numpy()
    assert rel_err(sr.detach().permute(0, 2, 3, 1).cpu().numpy(), ref_sr) < 1e-5
    loss = crit(sr, y, k)
    loss.backward()
    # the same weights through the reference-shaped trainer
    m2 = cuda_model(small_cfg, p, precision="tf32x3")
    L = pb.Losses((48, 48, 1))
    d = tempfile.mkdtemp(prefix="pv_")
    t = pb.ModelTrainer(m2, L.shiftCompensatedL1Loss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/ckpt", d + "/log")
    lossv, _ = t.forward_backward(lr, hr, mask)
    assert abs(float(loss.detach()) - lossv) <= 1e-6 * abs(lossv)
    assert torch.equal(net.theta.grad, t.grad_view())                                  # same kernels, fixed-order reductions: bit-identical
    # a torch optimizer updates the engine's arena in place, and the next forward sees it
    opt = torch.optim.SGD(net.parameters(), lr=1e-4)
    before = net.theta.detach().clone()
    opt.step()
    assert not torch.equal(before, net.theta.detach())
    assert torch.equal(net.theta.detach(), net.model.param_arena())
    sr2 = net(x)
    assert float((sr2 - sr).abs().max()) > 0
    names = [n for n, _ in net.named_variables()]
    assert names[0] == "mainConv1/v" and len(names) == 3 * (1 + 3 * small_cfg["numResBlocks"] + 3 + 1 + 3)


# ------------------------------------------------------------------------------------------------- 64 filters (BASELINE configs[4])
def test_64_filter_graph_with_sobel_l1_loss(small_cfg):
    """The scaled-up family of BASELINE configs[4] (num_filters = 64: expand to 512, decay to int(64 * 0.8) = 51 channels,
    modelsTF.py:177-189) with the Sobel + L1 loss (loss.py:86-97,219-224) on the dense fp32 engine: forward and every gradient."""
    import probav_b200 as pb
    import tempfile
    from probav_b200 import synth
    cfg = dict(small_cfg, numFilters=64, numResBlocks=2)
    om, p = oracle_and_params(cfg, seed=95)
    assert p["decConv_0/v"].shape[-1] == 51 and p["expConv_0/v"].shape[-1] == 512
    m = cuda_model(cfg, p, precision="fp32")
    lr, hr, mask = synth.make_batch(3, seed=96, hr_zero_under_mask=False)
    ol = OracleLosses((48, 48, 1))
    loss, g, sr, cps = loss_and_grads(om, ol, p, torch.from_numpy(lr).double(), torch.from_numpy(hr).double(), torch.from_numpy(mask), "sobel_l1_mix")
    assert rel_err(m(lr), sr.numpy()) < 1e-3
    L = pb.Losses((48, 48, 1))
    d = tempfile.mkdtemp(prefix="pv_")
    t = pb.ModelTrainer(m, L.shiftCompensatedL1EdgeLoss, L.shiftCompensatedcPSNR, pb.Nadam(5e-4), d + "/ckpt", d + "/log")
    lossv, psnrv = t.forward_backward(lr, hr, mask)
    assert abs(lossv - float(loss)) < 1e-3 * float(loss) and abs(psnrv - float(cps.mean())) < 0.01
    grads = t.get_grads()
    worst = max(rel_err(grads[k], v.numpy()) for k, v in g.items() if np.abs(v.numpy()).max() > 0)
    print(f"F=64, sobel_l1_mix: worst gradient rel err {worst:.2e}")
    assert worst < 4e-3          # 3-patch batch with a sign-based loss: see test_forward_backward_with_raw_hr


def test_cli_auto_precision_picks_the_engine_that_runs_the_graph():
    """train.py / test.py --precision auto: the error-compensated (train) or single-pass (predict) tensor-core engine for the graphs the row
    engine runs, the dense fp32 CUDA-core engine for the others (64 filters = BASELINE configs[4], 19 LR frames)."""
    import probav_b200 as pb
    from probav_b200 import cli
    from probav_b200.models import PRECISION
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = pb.parseConfig(os.path.join(root, "cfg", "p16t9c85r12.cfg"))
    base["num_res_blocks"] = 2
    for cfg_over, preferred, want in (({}, "tf32x3", "tf32x3"), ({}, "tf32", "tf32"), ({"num_filters": 64}, "tf32x3", "fp32"),
                                      ({"num_low_res_imgs": 19}, "tf32", "fp32")):
        m = cli._build_model(dict(base, **cfg_over), "NIR", "auto", preferred)
        assert m.cfg.precision == PRECISION[want], (cfg_over, preferred)
        m.close()
    m = cli._build_model(base, "NIR", "fp32_rows", "tf32x3")          # an explicit choice is taken as is
    assert m.cfg.precision == PRECISION["fp32_rows"]
    m.close()
