"""Shared by the parity tests: build the oracle and the CUDA model with identical synthetic weights."""
import numpy as np
import torch

from oracle.wdsr import OracleWDSR, init_params

NIR = (8075.2045, 3160.7272)


def oracle_and_params(cfg, seed=0, dtype=torch.float64, mean=NIR[0], std=NIR[1], maxShift=6):
    om = OracleWDSR(mean, std, maxShift, cfg["scale"], cfg["numFilters"], cfg["kernelSize"], cfg["numResBlocks"],
                    cfg["expRate"], cfg["decayRate"], cfg["numImgLR"], cfg["patchSizeLR"], cfg["isGrayScale"])
    p = init_params(om.specs, seed=seed, dtype=dtype)
    return om, p


def cuda_model(cfg, params, mean=NIR[0], std=NIR[1], maxShift=6, precision="fp32"):
    import probav_b200 as pb
    m = pb.WDSRConv3D("superResolutionNet", "NIR", mean, std, maxShift).build(**cfg, precision=precision)
    m.set_weights({k: v.detach().cpu().numpy().astype(np.float32) for k, v in params.items()})
    return m


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
