"""ctypes binding of libprobav_b200.so (C-ABI: include/probav_b200.h).

There is NO CPU fallback: if the shared library is missing, or no sm_100 device is present when a
compute entry point is called, this module raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libprobav_b200.so")

PV_LOSS = {"l1": 0, "l2": 1, "sobel_l1_mix": 2}
PV_OPT = {"sgd": 0, "adam": 1, "nadam": 2}


class pv_cfg(C.Structure):
    _fields_ = [("num_res_blocks", C.c_int32), ("num_low_res_imgs", C.c_int32), ("scale", C.c_int32),
                ("num_filters", C.c_int32), ("kernel_size", C.c_int32), ("exp_rate", C.c_int32),
                ("decay_rate", C.c_float), ("is_grayscale", C.c_int32), ("max_shift", C.c_int32),
                ("patch_size", C.c_int32), ("mean", C.c_float), ("std", C.c_float), ("precision", C.c_int32)]


class PvError(RuntimeError):
    pass


_P = C.c_void_p
_SIGS = {
    "pv_abi_version": (C.c_int, []),
    "pv_last_error": (C.c_char_p, []),
    "pv_device_count": (C.c_int, []),
    "pv_launch_count": (C.c_int64, []),
    "pv_model_create": (C.c_int, [C.POINTER(pv_cfg), C.c_int, C.POINTER(_P)]),
    "pv_model_destroy": (None, [_P]),
    "pv_model_param_count": (C.c_int, [_P]),
    "pv_model_param_info": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "pv_model_param_numel": (C.c_int64, [_P]),
    "pv_model_set_params": (C.c_int, [_P, _P, C.c_int64]),
    "pv_model_get_params": (C.c_int, [_P, _P, C.c_int64]),
    "pv_model_param_arena": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "pv_model_init_g_from_v": (C.c_int, [_P]),
    "pv_forward": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "pv_forward_host": (C.c_int, [_P, _P, C.c_int, _P]),
    "pv_resolve": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "pv_resolve_host": (C.c_int, [_P, _P, C.c_int, _P]),
    "pv_predict_scenes_host": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "pv_predict_from_scenes_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "pv_predict_from_scenes": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "pv_shift_loss": (C.c_int, [C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                _P, _P, _P, _P, _P, _P, _P, _P]),
    "pv_shift_loss_host": (C.c_int, [C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                     _P, _P, _P, _P, _P, _P, _P]),
    "pv_trainer_create": (C.c_int, [_P, C.c_int, C.c_float, C.c_int, C.POINTER(_P)]),
    "pv_trainer_destroy": (None, [_P]),
    "pv_train_step": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P]),
    "pv_train_step_host": (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    "pv_eval_step": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P]),
    "pv_eval_step_host": (C.c_int, [_P, _P, _P, _P, C.c_int, _P]),
    "pv_train_forward_backward": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_float, _P, _P]),
    "pv_train_forward_backward_staged": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_float, _P, C.c_int, C.POINTER(C.c_int64),
                                                   C.POINTER(C.c_int64), _P]),
    "pv_trainer_grad_arena": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "pv_apply_gradients": (C.c_int, [_P, _P]),
    "pv_trainer_get_state": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_double), _P, _P, C.c_int64]),
    "pv_trainer_set_state": (C.c_int, [_P, C.c_int64, C.c_double, _P, _P, C.c_int64]),
    "pv_trainer_set_lr": (C.c_int, [_P, C.c_float]),
    "pv_selftest": (C.c_int, [C.c_char_p, C.c_int]),
    "pv_trainer_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pv_trainer_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "pv_debug_read_buffer": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]),
    "pv_timing_enable": (C.c_int, [C.c_int]),
    "pv_timing_reset": (C.c_int, []),
    "pv_timing_report": (C.c_int, [C.c_char_p, C.c_int]),
}

_lib = None


def lib():
    """The loaded shared library; raises if it has not been built (python -m probav_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PvError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.pv_abi_version() != 1:
            raise PvError("libprobav_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(status: int):
    if status != 0:
        msg = lib().pv_last_error().decode(errors="replace")
        if status == -1:
            raise ValueError(msg)          # bad cfg: what Keras would raise while building the graph
        raise PvError(f"[pv_status {status}] {msg}")


def timing_report() -> dict:
    """{kernel class: dict(launches, ms, flops, bytes, exec_flops)} accumulated since pv_timing_reset (device-synchronising).
    flops / bytes are ALGORITHMIC (unpadded shapes, no recomputation); exec_flops is what the kernels execute."""
    n = lib().pv_timing_report(None, 0)
    buf = C.create_string_buffer(max(n, 1))
    lib().pv_timing_report(buf, n)
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms, fl, by, ex = line.split()
        out[name] = dict(launches=int(cnt), ms=float(ms), flops=float(fl), bytes=float(by), exec_flops=float(ex))
    return out


def selftest():
    """(number of failing tensor-core kernel configurations, report text)."""
    buf = C.create_string_buffer(8192)
    n = lib().pv_selftest(buf, 8192)
    return n, buf.value.decode()
