"""CPU: the oracle reproduces the committed golden vectors (tests/golden/make_golden.py), and the plain-C oracle
agrees with them too -- guards the checker against silent drift."""
import os

import numpy as np
import torch

from oracle import cshift
from oracle.losses import OracleLosses
from tests.helpers import oracle_and_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_shift_loss_golden_reproduced_by_both_oracles():
    z = np.load(os.path.join(GOLDEN, "shift_loss_golden.npz"))
    hr, sr, mask = (torch.from_numpy(z[k]) for k in ("hr", "sr", "mask"))
    L = OracleLosses((48, 48, 1))
    for kind, ck in (("l1", 0), ("l2", 1)):
        best, idx, cnt, stack = L.details(kind, hr.double(), mask, sr.double())
        assert np.array_equal(idx.numpy().astype(np.int32), z[f"best_shift_{kind}"])
        assert np.array_equal(cnt.numpy().astype(np.int32), z[f"clear_count_{kind}"])
        assert np.allclose(best.numpy(), z[f"loss_{kind}"], rtol=1e-12)
        sc, cn, _ = cshift.shift_scores(ck, z["hr"][..., 0], z["mask"][..., 0], z["sr"][..., 0])
        assert np.allclose(sc, z[f"stack_{kind}"], rtol=1e-9)
        assert np.array_equal(sc.argmin(1).astype(np.int32), z[f"best_shift_{kind}"])
    sc, _, _ = cshift.shift_scores(3, z["hr"][..., 0], z["mask"][..., 0], z["sr"][..., 0])
    assert np.allclose(sc.min(1), z["loss_l1edge"], rtol=1e-9)
    assert np.array_equal(sc.argmin(1).astype(np.int32), z["best_shift_l1edge"])
    assert np.allclose(L.shiftCompensatedcPSNR(hr.double(), mask, sr.double()).numpy(), z["cpsnr"], atol=1e-9)


def test_wdsr_golden_reproduced(small_cfg):
    z = np.load(os.path.join(GOLDEN, "wdsr_small_golden.npz"))
    om, p = oracle_and_params(small_cfg, seed=int(z["weight_seed"]))
    with torch.no_grad():
        sr = om.forward(p, torch.from_numpy(z["lr"]).double()).numpy()
    assert np.allclose(sr, z["sr"], rtol=1e-12, atol=1e-9)
    # an fp32 run of the same oracle stays inside the GPU tolerance budget (sizes the 1e-3 bar)
    p32 = {k: v.float() for k, v in p.items()}
    with torch.no_grad():
        sr32 = om.forward(p32, torch.from_numpy(z["lr"])).numpy()
    assert np.abs(sr32 - z["sr"]).max() / np.abs(z["sr"]).max() < 1e-4
