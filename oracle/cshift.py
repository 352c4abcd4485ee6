"""ctypes loader for oracle/shift_loss.c (test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_shift_loss.so")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib():
    if not os.path.exists(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    lib.oracle_shift_scores.restype = None
    return lib


def shift_scores(kind: int, hr: np.ndarray, mask: np.ndarray, sr: np.ndarray, border: int = 3):
    """hr, sr [B,H,W] float32; mask [B,H,W] bool/uint8 -> (scores, counts, biases) each [B, S*S] float64."""
    hr = np.ascontiguousarray(hr, dtype=np.float32)
    sr = np.ascontiguousarray(sr, dtype=np.float32)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    B, H, W = hr.shape
    S = 2 * border + 1
    sc = np.empty((B, S * S), np.float64)
    cn = np.empty((B, S * S), np.float64)
    bi = np.empty((B, S * S), np.float64)
    P = ctypes.c_void_p
    _lib().oracle_shift_scores(ctypes.c_int(kind), P(hr.ctypes.data), P(mask.ctypes.data), P(sr.ctypes.data),
                               ctypes.c_int(B), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(border),
                               P(sc.ctypes.data), P(cn.ctypes.data), P(bi.ctypes.data))
    return sc, cn, bi
