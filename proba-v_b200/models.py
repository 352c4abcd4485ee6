"""Graph builder with the reference's surface: WDSRConv3D(name, band, mean, std, maxShift).build(...)
(reference models/modelsTF.py:8-43).  The returned WDSRModel stands in for the tf.keras.Model: callable
as model(lr_batch, training=False) -> [B, scale*patch, scale*patch, 1], exposes trainable_variables,
get_weights / set_weights, count_params.  All compute happens in libprobav_b200.so (sm_100a kernels)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List

import numpy as np

from . import _buf
from ._lib import check, lib, pv_cfg

PRECISION = {"fp32": 0, "tf32": 1, "fp32_rows": 3, "tf32x3": 4}


class Variable:
    """Read/write view of one weight tensor (TF layout) inside the model's flat arena."""

    def __init__(self, model: "WDSRModel", name: str, shape, offset: int, numel: int):
        self._model, self.name, self.shape, self.offset, self.numel = model, name, tuple(shape), offset, numel

    def numpy(self) -> np.ndarray:
        return self._model.get_flat()[self.offset:self.offset + self.numel].reshape(self.shape).copy()

    def assign(self, value):
        flat = self._model.get_flat()
        flat[self.offset:self.offset + self.numel] = np.asarray(value, np.float32).reshape(-1)
        self._model.set_flat(flat)

    def __repr__(self):
        return f"<Variable {self.name} shape={self.shape}>"


class WDSRModel:
    def __init__(self, name: str, cfg: pv_cfg, device: int):
        self.name = name
        self.cfg = cfg
        self.device = device
        h = C.c_void_p()
        check(lib().pv_model_create(C.byref(cfg), device, C.byref(h)))
        self._h = h
        self.S = cfg.patch_size + cfg.max_shift
        self.T = cfg.num_low_res_imgs
        self.out_side = cfg.patch_size * cfg.scale
        self.input_shape = (None, self.S, self.S, self.T, 1)
        self.output_shape = (None, self.out_side, self.out_side, 1)
        self._vars: List[Variable] = []
        n = lib().pv_model_param_count(self._h)
        for i in range(n):
            nm = C.create_string_buffer(128)
            rank, off, numel = C.c_int(), C.c_int64(), C.c_int64()
            shape = (C.c_int64 * 5)()
            check(lib().pv_model_param_info(self._h, i, nm, 128, C.byref(rank), shape, C.byref(off), C.byref(numel)))
            self._vars.append(Variable(self, nm.value.decode(), [shape[k] for k in range(rank.value)], off.value, numel.value))
        self.nparams = int(lib().pv_model_param_numel(self._h))

    # ---- Keras-like surface
    @property
    def trainable_variables(self) -> List[Variable]:
        return list(self._vars)

    def count_params(self) -> int:
        return self.nparams

    def get_flat(self) -> np.ndarray:
        out = np.empty(self.nparams, np.float32)
        check(lib().pv_model_get_params(self._h, _buf.ptr(out), self.nparams))
        return out

    def set_flat(self, flat):
        flat = np.ascontiguousarray(flat, np.float32).reshape(-1)
        check(lib().pv_model_set_params(self._h, _buf.ptr(flat), flat.size))

    def get_weights(self) -> Dict[str, np.ndarray]:
        flat = self.get_flat()
        return {v.name: flat[v.offset:v.offset + v.numel].reshape(v.shape).copy() for v in self._vars}

    def set_weights(self, weights: Dict[str, np.ndarray]):
        flat = self.get_flat()
        for v in self._vars:
            if v.name in weights:
                w = np.asarray(weights[v.name], np.float32)
                if w.size != v.numel:
                    raise ValueError(f"{v.name}: expected shape {v.shape}, got {w.shape}")
                flat[v.offset:v.offset + v.numel] = w.reshape(-1)
        self.set_flat(flat)

    def layer_names(self) -> List[str]:
        seen = []
        for v in self._vars:
            layer = v.name.rsplit("/", 1)[0]
            if layer not in seen:
                seen.append(layer)
        return seen

    def restore_checkpoint(self, ckpt_dir_or_prefix: str) -> dict:
        """tf.train.Checkpoint(model=...).restore(manager.latest_checkpoint) of test.py:58-67: loads the weights of a TensorFlow
        checkpoint (written by the reference or by ModelTrainer.save) and returns its {"step", "psnr", "save_counter"}."""
        import os
        from . import tfckpt
        prefix = ckpt_dir_or_prefix
        if os.path.isdir(prefix):
            prefix = tfckpt.latest_checkpoint(prefix)
            if prefix is None:
                raise FileNotFoundError(f"no checkpoint in {ckpt_dir_or_prefix}")
        z = tfckpt.load_checkpoint(prefix, self.layer_names())
        self.set_weights(z["weights"])
        return {k: z[k] for k in ("step", "psnr", "save_counter")}

    def init_weights(self, seed: int = 0):
        """Keras defaults under TFA WeightNormalization(data_init=False): v Glorot-uniform, bias 0, g <- ||v||."""
        rng = np.random.default_rng(seed)
        flat = np.zeros(self.nparams, np.float32)
        for v in self._vars:
            if v.name.endswith("/v"):
                rf = math.prod(v.shape[:-2])
                limit = math.sqrt(6.0 / (rf * v.shape[-2] + rf * v.shape[-1]))
                flat[v.offset:v.offset + v.numel] = rng.uniform(-limit, limit, v.numel)
        self.set_flat(flat)
        check(lib().pv_model_init_g_from_v(self._h))

    def param_arena(self):
        """torch view (no copy) of the flat device arena -- used for the data-parallel weight broadcast."""
        p, n = C.c_void_p(), C.c_int64()
        check(lib().pv_model_param_arena(self._h, C.byref(p), C.byref(n)))
        return _buf.view_device_floats(p.value, n.value, f"cuda:{self.device}")

    def __call__(self, lr_batch, training: bool = False, resolve: bool = False):
        """model(lr_batch): [B,S,S,T,1] -> [B,sP,sP,1].  numpy in -> numpy out; torch CUDA in -> torch CUDA out."""
        shp = tuple(lr_batch.shape)
        if len(shp) != 5 or shp[1:] != (self.S, self.S, self.T, 1):
            raise ValueError(f"expected LR batch [B,{self.S},{self.S},{self.T},1], got {shp}")
        B = shp[0]
        if _buf.is_cuda_tensor(lr_batch):
            import torch
            x = _buf.dev_tensor(lr_batch, torch.float32, lr_batch.device)
            y = torch.empty((B, self.out_side, self.out_side, 1), dtype=torch.float32, device=x.device)
            fn = lib().pv_resolve if resolve else lib().pv_forward
            check(fn(self._h, _buf.ptr(x), B, _buf.ptr(y), _buf.current_stream_ptr(x.device)))
            return y
        x = _buf.host_array(lr_batch, np.float32)
        y = np.empty((B, self.out_side, self.out_side, 1), np.float32)
        fn = lib().pv_resolve_host if resolve else lib().pv_forward_host
        check(fn(self._h, _buf.ptr(x), B, _buf.ptr(y)))
        return y

    def predict_scenes(self, patchLR) -> np.ndarray:
        """[nscenes, n*n, S,S,T,1] host patches -> [nscenes, n*P, n*P, 1] stitched, resolved (test.py:103-160)."""
        x = _buf.host_array(patchLR, np.float32)
        ns, pps = x.shape[0], x.shape[1]
        n = int(round(pps ** 0.5))
        out = np.empty((ns, n * self.out_side, n * self.out_side, 1), np.float32)
        check(lib().pv_predict_scenes_host(self._h, _buf.ptr(x), ns, pps, _buf.ptr(out)))
        return out

    def predict_from_scenes(self, scenesLR) -> np.ndarray:
        """[nscenes, T, H, W] LR scenes -> [nscenes, s*H, s*W, 1]; patching + stitching on device.  numpy in -> numpy out
        (pinned double-buffered transfers overlapped with the forward pass); torch CUDA in -> torch CUDA out (asynchronous)."""
        if _buf.is_cuda_tensor(scenesLR):
            import torch
            x = _buf.dev_tensor(scenesLR, torch.float32, scenesLR.device)
            ns, T, H, W = x.shape
            if T != self.T:
                raise ValueError(f"expected {self.T} LR frames, got {T}")
            s = self.cfg.scale
            y = torch.empty((ns, s * H, s * W, 1), dtype=torch.float32, device=x.device)
            check(lib().pv_predict_from_scenes(self._h, _buf.ptr(x), ns, H, W, _buf.ptr(y), _buf.current_stream_ptr(x.device)))
            return y
        x = _buf.host_array(scenesLR, np.float32)
        ns, T, H, W = x.shape
        if T != self.T:
            raise ValueError(f"expected {self.T} LR frames, got {T}")
        s = self.cfg.scale
        out = np.empty((ns, s * H, s * W, 1), np.float32)
        check(lib().pv_predict_from_scenes_host(self._h, _buf.ptr(x), ns, H, W, _buf.ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().pv_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WDSRConv3D:
    """Same constructor and build() arguments as the reference class (modelsTF.py:8-17)."""

    def __init__(self, name, band, mean, std, maxShift):
        self.name, self.band, self.mean, self.std, self.maxShift = name, band, mean, std, maxShift

    def build(self, scale: int, numFilters: int, kernelSize: tuple, numResBlocks: int, expRate: int,
              decayRate: float, numImgLR: int, patchSizeLR: int, isGrayScale: bool,
              precision: str = "fp32", device: int = None, seed: int = 0) -> WDSRModel:
        ks = kernelSize if isinstance(kernelSize, int) else kernelSize[0]
        if not isinstance(kernelSize, int) and any(k != ks for k in kernelSize):
            raise ValueError(f"anisotropic kernelSize {kernelSize} is not supported")
        if device is None:
            device = _default_device()
        cfg = pv_cfg(num_res_blocks=numResBlocks, num_low_res_imgs=numImgLR, scale=scale, num_filters=numFilters,
                     kernel_size=ks, exp_rate=expRate, decay_rate=decayRate, is_grayscale=int(bool(isGrayScale)),
                     max_shift=self.maxShift, patch_size=patchSizeLR, mean=self.mean, std=self.std,
                     precision=PRECISION[precision])
        model = WDSRModel(f"WDSRConv3D_{self.band}_{self.name}", cfg, device)
        model.init_weights(seed)
        return model


def _default_device() -> int:
    import os
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def build_from_config(config: dict, band: str = "NIR", name: str = "superResolutionNet", **kw) -> WDSRModel:
    """train.py:47-52,66-74 / test.py:40-56: per-band constants + WDSRConv3D(...).build(**cfg)."""
    from .synth import BAND_STATS
    mean, std = BAND_STATS["NIR" if band == "NIR" else "RED"]
    k = config["kernel_size"]
    return WDSRConv3D(name=name, band=band, mean=mean, std=std, maxShift=config["max_shift"]).build(
        scale=config["scale"], numFilters=config["num_filters"], kernelSize=(k, k, k),
        numResBlocks=config["num_res_blocks"], expRate=config["exp_rate"], decayRate=config["decay_rate"],
        numImgLR=config["num_low_res_imgs"], patchSizeLR=config["patch_size"], isGrayScale=config["is_grayscale"], **kw)
