"""Host/device buffer helpers: the C-ABI takes raw pointers; Python callers hand numpy arrays (host) or
torch CUDA tensors (device)."""
from __future__ import annotations

import ctypes as C

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def is_cuda_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


def host_array(x, dtype) -> np.ndarray:
    if torch is not None and isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def dev_tensor(x, dtype, device):
    """contiguous torch CUDA tensor of `dtype` on `device` (copies only when needed; pinned host tensors copy async)."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    if x.dtype == torch.bool:
        x = x.contiguous().view(torch.uint8)          # bool and uint8 share the byte layout (True = 1)
    return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()


def ptr(x) -> C.c_void_p:
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(x.data_ptr())


def current_stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _RawDeviceArray:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can view it (no copy)."""

    def __init__(self, p: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (p, False), "version": 2}


def view_device_floats(p: int, n: int, device):
    return torch.as_tensor(_RawDeviceArray(p, n), device=device)
