"""Step loops with the reference's surface (reference models/trainClass.py:17-143):
ModelTrainer(model, loss, metric, optimizer, ckptDir, logDir, multiGPU=True, evalStep=1000) with
.fitTrainData / .trainStep / .testStep / .restore / .model.  One trainStep = pv_train_step (fwd, shift loss +
cPSNR in one pass, fused loss backward, conv backward, weight-norm backward, fused optimizer); under
torch.distributed the gradient arena is all-reduced once between backward and apply_gradients."""
from __future__ import annotations

import ctypes as C
import glob
import logging
import os
import time
from typing import List

import numpy as np

from . import _buf, parallel, tbevents, tfckpt
from ._lib import PV_LOSS, PV_OPT, check, lib
from .loss import LOSS_KIND_OF_METHOD

logging.basicConfig(format="%(asctime)s - %(message)s", level=logging.INFO)
logger = logging.getLogger("probav_b200")


class _nvtx:
    """NVTX range around a step when PV_NVTX=1 (torch.cuda.nvtx; a no-op otherwise)."""
    _on = os.environ.get("PV_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if self._on:
            import torch
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if self._on:
            import torch
            torch.cuda.nvtx.range_pop()


class Mean:
    """tf.keras.metrics.Mean: running mean, count-weighted by update calls (trainClass.py:43-46)."""

    def __init__(self, name=""):
        self.name = name
        self.reset_states()

    def reset_states(self):
        self.total, self.count = 0.0, 0

    def __call__(self, value, weight=1):
        self.total += float(value) * weight
        self.count += weight

    def result(self) -> float:
        return self.total / self.count if self.count else 0.0


def shuffled_index_stream(n: int, epochs: int, buffer_size: int, rng: np.random.Generator):
    """tf.data shuffle(buffer).repeat(epochs): a streaming shuffle buffer refilled in dataset order, reshuffled each epoch."""
    for _ in range(epochs):
        buf = list(range(min(buffer_size, n)))
        nxt = len(buf)
        while buf:
            k = int(rng.integers(len(buf)))
            yield buf[k]
            if nxt < n:
                buf[k] = nxt
                nxt += 1
            else:
                buf[k] = buf[-1]
                buf.pop()


def batched(stream, batch_size: int):
    """.batch(B) without drop_remainder: the final batch may be partial (utils/utils.py:32-34)."""
    cur = []
    for i in stream:
        cur.append(i)
        if len(cur) == batch_size:
            yield np.asarray(cur)
            cur = []
    if cur:
        yield np.asarray(cur)


class ModelTrainer:
    def __init__(self, model, loss, metric, optimizer, ckptDir, logDir, multiGPU=True, evalStep=1000, seed=0):
        os.makedirs(ckptDir, exist_ok=True)
        os.makedirs(logDir, exist_ok=True)
        kind = LOSS_KIND_OF_METHOD.get(getattr(loss, "__name__", ""), None)
        if kind is None:
            raise ValueError("loss must be one of Losses.shiftCompensated{L1Loss,L2Loss,L1EdgeLoss}")
        if getattr(metric, "__name__", "") != "shiftCompensatedcPSNR":
            raise ValueError("metric must be Losses.shiftCompensatedcPSNR (train.py:104)")
        self._model = model
        self.loss, self.metric, self.optimizer = loss, metric, optimizer
        self.loss_kind = kind
        self.ckptDir, self.logDir = ckptDir, logDir
        self.trainLoss, self.trainPSNR = Mean("trainLoss"), Mean("trainPSNR")
        self.testLoss, self.testPSNR = Mean("testLoss"), Mean("testPSNR")
        self.evalStep = evalStep
        self.multiGPU = multiGPU
        self.strategy = None
        self.step = 0               # ckpt.step
        self.psnr = 1.0             # ckpt.psnr
        self.max_to_keep = 5
        self.save_counter = 0       # ckpt.save_counter
        self._rng = np.random.default_rng(seed)
        self._val_rng = np.random.default_rng(seed + 1)     # the training index stream is drawn from the prefetch thread
        h = C.c_void_p()
        check(lib().pv_trainer_create(model._h, PV_OPT[optimizer.kind], optimizer.learning_rate, PV_LOSS[kind], C.byref(h)))
        self._h = h
        self._out = np.zeros(2, np.float32)
        # tf.summary.create_file_writer(logDir) (trainClass.py:41): a TensorBoard event file, rank 0 only
        self._scalars = tbevents.SummaryWriter(logDir) if parallel.world()[0] == 0 else None
        self._grad_view = None
        self.restore()

    @property
    def model(self):
        return self._model

    # ---- checkpoint: tf.train.Checkpoint(step, psnr, optimizer, model) + CheckpointManager(max_to_keep=5) (trainClass.py:33-39),
    # written in TensorFlow's own tensor-bundle format (tfckpt.py): ckpt-N.index, ckpt-N.data-00000-of-00001, `checkpoint`
    def _layer_names(self):
        return self._model.layer_names()

    def _legacy_latest(self):
        files = sorted(glob.glob(os.path.join(self.ckptDir, "ckpt-*.npz")), key=lambda f: int(f.split("-")[-1][:-4]))
        return files[-1] if files else None

    def _split(self, flat):
        return {v.name: flat[v.offset:v.offset + v.numel].reshape(v.shape) for v in self._model.trainable_variables}

    def _join(self, d):
        flat = np.zeros(self._model.nparams, np.float32)
        for v in self._model.trainable_variables:
            if v.name in d:
                flat[v.offset:v.offset + v.numel] = np.asarray(d[v.name], np.float32).reshape(-1)
        return flat

    def restore(self):
        """trainClass.py:52-59: restore the newest checkpoint of ckptDir, if any (TF bundle; .npz of earlier versions too)."""
        prefix = tfckpt.latest_checkpoint(self.ckptDir)
        n = self._model.nparams
        if prefix:
            z = tfckpt.load_checkpoint(prefix, self._layer_names())
            self._model.set_weights(z["weights"])
            o = z["opt"]
            if o is not None and o.get("m") and o.get("v"):
                m1, m2 = self._join(o["m"]), self._join(o["v"])
                check(lib().pv_trainer_set_state(self._h, int(o["iter"]), float(o.get("momentum_cache", 1.0)), _buf.ptr(m1), _buf.ptr(m2), n))
            self.step = int(z["step"]) if z["step"] is not None else 0
            self.psnr = float(z["psnr"]) if z["psnr"] is not None else self.psnr
            self.save_counter = int(z["save_counter"] or 0)
        else:
            f = self._legacy_latest()
            if not f:
                return
            z = np.load(f)
            self._model.set_flat(z["params"])
            m1 = np.ascontiguousarray(z["opt_m"], np.float32)
            m2 = np.ascontiguousarray(z["opt_v"], np.float32)
            check(lib().pv_trainer_set_state(self._h, int(z["opt_iter"]), float(z["momentum_cache"]), _buf.ptr(m1), _buf.ptr(m2), n))
            self.step, self.psnr = int(z["step"]), float(z["psnr"])
        print(f"[ INFO ] Model restored from checkpoint at step {self.step}.")

    def save(self):
        """CheckpointManager.save (trainClass.py:118-120): ckpt-<save_counter>, keeps the newest max_to_keep."""
        if parallel.world()[0] != 0:
            return None
        n = self._model.nparams
        it, mc = C.c_int64(), C.c_double()
        m1, m2 = np.empty(n, np.float32), np.empty(n, np.float32)
        check(lib().pv_trainer_get_state(self._h, C.byref(it), C.byref(mc), _buf.ptr(m1), _buf.ptr(m2), n))
        self.save_counter += 1
        name = f"ckpt-{self.save_counter}"
        prefix = os.path.join(self.ckptDir, name)
        o = self.optimizer
        opt = {"iter": it.value, "learning_rate": o.learning_rate, "beta_1": getattr(o, "beta_1", 0.9), "beta_2": getattr(o, "beta_2", 0.999),
               "decay": 0.0, "momentum_cache": mc.value, "m": self._split(m1), "v": self._split(m2)}
        tfckpt.save_checkpoint(prefix, self._layer_names(), self._model.get_weights(), self.step, self.psnr, self.save_counter, opt)
        st = tfckpt.read_checkpoint_state(self.ckptDir)
        paths, stamps = st["all_model_checkpoint_paths"] + [name], st["all_model_checkpoint_timestamps"] + [time.time()]
        while len(paths) > self.max_to_keep:
            old = os.path.join(self.ckptDir, paths.pop(0))
            stamps.pop(0)
            for f in glob.glob(old + ".index") + glob.glob(old + ".data-*"):
                os.remove(f)
        tfckpt.write_checkpoint_state(self.ckptDir, paths, stamps, st["last_preserved_timestamp"])
        return prefix + ".index"

    # ---- steps
    def _run(self, fn_host, fn_dev, patchLR, patchHR, maskHR):
        B = int(patchLR.shape[0])
        if _buf.is_cuda_tensor(patchLR):
            import torch
            dev = patchLR.device
            x = _buf.dev_tensor(patchLR, torch.float32, dev)
            y = _buf.dev_tensor(patchHR, torch.float32, dev)
            m = _buf.dev_tensor(maskHR, torch.uint8, dev)
            out = torch.empty(2, dtype=torch.float32, device=dev)
            check(fn_dev(self._h, _buf.ptr(x), _buf.ptr(y), _buf.ptr(m), B, _buf.ptr(out), _buf.current_stream_ptr(dev)))
            return out
        x = _buf.host_array(patchLR, np.float32)
        y = _buf.host_array(patchHR, np.float32)
        m = _buf.host_array(maskHR, np.uint8)
        check(fn_host(self._h, _buf.ptr(x), _buf.ptr(y), _buf.ptr(m), B, _buf.ptr(self._out)))
        return self._out.copy()

    def trainStep(self, patchLR, patchHR, maskHR, global_batch: int = None, sync: bool = True):
        """trainClass.py:124-135.  Returns (loss, mean cPSNR) of the global batch and updates the running means.
        sync=False (device tensors only) skips the per-step host read the reference's logging forces and returns the
        device tensor [loss, cPSNR] of this rank's shard instead."""
        rank, ws = parallel.world()
        with _nvtx("pv.trainStep"):       # PV_NVTX=1: NVTX ranges for Nsight timelines (tracing)
            if ws == 1:
                out = self._run(lib().pv_train_step_host, lib().pv_train_step, patchLR, patchHR, maskHR)
            else:
                out = self._dp_step(patchLR, patchHR, maskHR, global_batch)
        if not sync and _buf.is_cuda_tensor(out):
            return out
        lossv, psnrv = float(out[0]), float(out[1])
        if not sync:            # host arrays in: the C-ABI host entry already synchronised; the fit loop folds the values itself
            if ws > 1:
                lossv, psnrv = parallel.reduce_metrics(lossv, psnrv, int(patchLR.shape[0]), device=getattr(out, "device", None))
            self._last_host = (lossv, psnrv)
            return None
        if ws > 1:
            lossv, psnrv = parallel.reduce_metrics(lossv, psnrv, int(patchLR.shape[0]), device=out.device)
        self.trainLoss(lossv)
        self.trainPSNR(psnrv)
        return lossv, psnrv

    def _dp_step(self, patchLR, patchHR, maskHR, global_batch):
        """fwd/bwd on the local shard in two stages; the all-reduce(SUM) of the first gradient bucket (tail, reducers, last
        R/2 blocks) runs on NCCL's stream while the second half of the backward pass computes; identical update on every
        rank.  The two buckets are contiguous ranges of the flat gradient arena (never per-variable messages)."""
        import torch
        dev = patchLR.device if _buf.is_cuda_tensor(patchLR) else torch.device(f"cuda:{self._model.device}")
        B = int(patchLR.shape[0])
        ws = parallel.world()[1]
        gb = global_batch if global_batch else B * ws
        x = _buf.dev_tensor(patchLR, torch.float32, dev)
        y = _buf.dev_tensor(patchHR, torch.float32, dev)
        m = _buf.dev_tensor(maskHR, torch.uint8, dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        st = _buf.current_stream_ptr(dev)
        g = self.grad_view()
        if os.environ.get("PV_DP_OVERLAP", "1") == "0":       # one all-reduce of the whole arena after backward (A/B measurements)
            check(lib().pv_train_forward_backward(self._h, _buf.ptr(x), _buf.ptr(y), _buf.ptr(m), B, parallel.grad_scale(gb), _buf.ptr(out), st))
            parallel.allreduce_sum_(g)
            check(lib().pv_apply_gradients(self._h, st))
            return out
        lo, hi = C.c_int64(), C.c_int64()
        works = []
        for stage in (0, 1):
            check(lib().pv_train_forward_backward_staged(self._h, _buf.ptr(x), _buf.ptr(y), _buf.ptr(m), B, parallel.grad_scale(gb),
                                                         _buf.ptr(out), stage, C.byref(lo), C.byref(hi), st))
            if hi.value > lo.value:
                works.append(parallel.allreduce_sum_async_(g[lo.value:hi.value]))
        for w in works:
            if w is not None:
                w.wait()                      # stream-level wait: the optimizer kernel is ordered after both collectives
        check(lib().pv_apply_gradients(self._h, st))
        return out

    def testStep(self, patchLR, patchHR, maskHR):
        """trainClass.py:137-143."""
        with _nvtx("pv.testStep"):
            out = self._run(lib().pv_eval_step_host, lib().pv_eval_step, patchLR, patchHR, maskHR)
        lossv, psnrv = float(out[0]), float(out[1])
        self.testLoss(lossv)
        self.testPSNR(psnrv)
        return lossv, psnrv

    # ---- introspection used by the parity tests and the DP path
    def forward_backward(self, patchLR, patchHR, maskHR, grad_scale: float = None):
        """tape.gradient only (no optimizer step): returns (loss, mean cPSNR); gradients via get_grads()."""
        import torch
        dev = torch.device(f"cuda:{self._model.device}")
        B = int(patchLR.shape[0])
        x = _buf.dev_tensor(patchLR, torch.float32, dev)
        y = _buf.dev_tensor(patchHR, torch.float32, dev)
        m = _buf.dev_tensor(maskHR, torch.uint8, dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        gs = (1.0 / B) if grad_scale is None else float(grad_scale)
        check(lib().pv_train_forward_backward(self._h, _buf.ptr(x), _buf.ptr(y), _buf.ptr(m), B, gs, _buf.ptr(out),
                                              _buf.current_stream_ptr(dev)))
        o = out.cpu().numpy()
        return float(o[0]), float(o[1])

    def grad_view(self):
        """torch view (no copy) of the flat gradient arena (same layout as the parameter arena)."""
        if self._grad_view is None:
            p, n = C.c_void_p(), C.c_int64()
            check(lib().pv_trainer_grad_arena(self._h, C.byref(p), C.byref(n)))
            self._grad_view = _buf.view_device_floats(p.value, n.value, f"cuda:{self._model.device}")
        return self._grad_view

    def get_grads(self) -> dict:
        flat = self.grad_view().cpu().numpy()
        return {v.name: flat[v.offset:v.offset + v.numel].reshape(v.shape).copy() for v in self._model.trainable_variables}

    def apply_gradients(self):
        import torch
        check(lib().pv_apply_gradients(self._h, _buf.current_stream_ptr(torch.device(f"cuda:{self._model.device}"))))

    def _scalar(self, tag, value, step):
        if self._scalars:
            self._scalars.scalar(tag, value, step)

    # ---- fit loop
    def fitTrainData(self, X, y, globalBatchSize: int, epochs: int, valData: List, bufferSize: int = 256,
                     valSteps: int = 64, saveBestOnly: bool = True, initEpoch: int = 0, maxSteps: int = None,
                     logEvery: int = 1, prefetch: bool = True):
        """trainClass.py:61-122.  X [N,S,S,T,1]; y = [HR [N,..,1], mask]; valData = [X_val, y_val, mask_val].
        Under torch.distributed every rank walks the same index stream and takes its shard of each global batch."""
        rank, ws = parallel.world()
        yHR, yMask = y
        n = len(X)
        totalSteps = int(n / globalBatchSize)
        globalStep = self.step
        step = globalStep % totalSteps if totalSteps else 0
        epoch = initEpoch
        stream = batched(shuffled_index_stream(n, epochs, bufferSize, self._rng), globalBatchSize)
        if prefetch:      # .prefetch(AUTOTUNE) of utils/utils.py:32-34: this rank's shard is gathered into pinned memory and copied
            import torch  # to the GPU on a side stream while the previous step computes (pipeline.py)
            from .pipeline import PrefetchLoader
            batches = PrefetchLoader((X, yHR, yMask), stream, device=torch.device(f"cuda:{self._model.device}"), rank=rank,
                                     world_size=ws, max_batch=globalBatchSize)
        else:
            def host_batches():
                for idx in stream:
                    lo, hi = parallel.shard_bounds(len(idx), rank, ws)
                    sel = np.sort(idx[lo:hi]) if hi > lo else idx[:0]
                    yield len(idx), (X[sel], yHR[sel], yMask[sel])
            batches = host_batches()
        logger.info("[ INFO ] Begin training...")
        done = 0
        # The reference reads the running means back on every step (logger + tf.summary, trainClass.py:96-102), which would drain
        # the GPU once per step.  Here the step's [loss, cPSNR] (summed over ranks on the device) is copied to pinned memory
        # asynchronously and folded into the means -- and logged, with the same values and step numbers -- one iteration later,
        # after the NEXT step has been queued; evaluation, epoch boundaries and the end of the loop flush it first.
        pending = None

        def finish(p):
            if p is None:
                return None
            ev, host, n_glob, p_step, p_gstep, p_epoch = p
            if ev is not None:
                ev.synchronize()
            lossv, psnrv = float(host[0]) / n_glob, float(host[1]) / n_glob
            self.trainLoss(lossv)
            self.trainPSNR(psnrv)
            if logEvery and (p_step % logEvery == 0) and rank == 0:
                logger.info(f"[ EPOCH {p_epoch}/{epochs} ] - [ STEP {p_step}/{totalSteps} ] Loss: {self.trainLoss.result():.6f}, "
                            f"cPSNR: {self.trainPSNR.result():.3f}")
            self._scalar("Train PSNR", self.trainPSNR.result(), p_gstep)
            self._scalar("Train loss", self.trainLoss.result(), p_gstep)
            return None

        for gb, (xb, yb, mb) in batches:
            if totalSteps - step == 0:
                pending = finish(pending)
                epoch += 1
                step = self.step % totalSteps
                logger.info(f"[ ***************  NEW EPOCH  *************** ] Epoch number {epoch}")
                for mtr in (self.trainLoss, self.trainPSNR, self.testLoss, self.testPSNR):
                    mtr.reset_states()
            step += 1
            globalStep += 1
            out = self.trainStep(xb, yb, mb, global_batch=gb, sync=False)
            self.step += 1
            if _buf.is_cuda_tensor(out):
                cur = (*self._metrics_async(out, int(xb.shape[0]), gb), step, globalStep, epoch)
                finish(pending)                    # the previous step's numbers, while this step runs
                pending = cur
            else:                                  # host path (prefetch=False with numpy batches): trainStep already synchronised
                pending = finish(pending)
                finish((None, [self._last_host[0], self._last_host[1]], 1, step, globalStep, epoch))
            if step != 0 and (step % self.evalStep) == 0:
                pending = finish(pending)
                self.testLoss.reset_states()
                self.testPSNR.reset_states()
                Xv, yv, mv = valData
                vstream = batched(shuffled_index_stream(len(Xv), 1, bufferSize, self._val_rng), globalBatchSize)
                for k, vidx in enumerate(vstream):
                    if k >= valSteps:
                        break
                    vsel = np.sort(vidx)
                    if ws == 1:
                        self.testStep(Xv[vsel], yv[vsel], mv[vsel])
                    else:
                        # data parallel: every rank scores its shard of the validation batch and the sums are all-reduced (each rank
                        # used to evaluate the whole global batch: world-size times the work and the activation memory)
                        lo, hi = parallel.shard_bounds(len(vsel), rank, ws)
                        sub = vsel[lo:hi]
                        lv = pv_ = 0.0
                        if len(sub):
                            o = self._run(lib().pv_eval_step_host, lib().pv_eval_step, Xv[sub], yv[sub], mv[sub])
                            lv, pv_ = float(o[0]), float(o[1])
                        import torch
                        gl, gp = parallel.reduce_metrics(lv, pv_, len(sub), device=torch.device(f"cuda:{self._model.device}"))
                        self.testLoss(gl)
                        self.testPSNR(gp)
                self._scalar("Test loss", self.testLoss.result(), globalStep)
                self._scalar("Test PSNR", self.testPSNR.result(), globalStep)
                if rank == 0:
                    logger.info(f"[ *************** VAL INFO *************** ] Validation Loss: {self.testLoss.result():.6f}, "
                                f"Validation PSNR: {self.testPSNR.result():.3f}")
                if self._scalars:
                    self._scalars.flush()
                if not (saveBestOnly and self.testPSNR.result() <= self.psnr):
                    logger.info("[ SAVE ] Saving checkpoint...")
                    self.psnr = self.testPSNR.result()
                    self.save()
            done += 1
            if maxSteps and done >= maxSteps:
                break
        pending = finish(pending)
        if self._scalars:
            self._scalars.flush()

    def _metrics_async(self, out_dev, n_local: int, n_global: int):
        """[loss, cPSNR] of this rank's shard (device) -> pinned host [sum loss, sum cPSNR] over the GLOBAL batch, asynchronously:
        returns (event, pinned tensor, n_global).  Two pinned slots alternate, so a slot is read before it is reused."""
        import torch
        if not hasattr(self, "_mslots"):
            self._mslots = [torch.empty(2, dtype=torch.float64).pin_memory() for _ in range(2)]
            self._mslot = 0
        v = out_dev.double() * float(n_local)
        if parallel.world()[1] > 1:
            parallel.allreduce_sum_(v)
        host = self._mslots[self._mslot]
        self._mslot ^= 1
        host.copy_(v, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(out_dev.device))
        return ev, host, n_global

    def close(self):
        if getattr(self, "_h", None):
            lib().pv_trainer_destroy(self._h)
            self._h = None
        if getattr(self, "_scalars", None):
            self._scalars.close()
            self._scalars = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
