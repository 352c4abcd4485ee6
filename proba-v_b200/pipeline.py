"""Input pipeline of the fit loop: tf.data's shuffle(buffer).repeat(epochs).batch(B).prefetch(AUTOTUNE)
(reference utils/utils.py:32-39, used at models/trainClass.py:69-71) re-designed for one process per GPU.

The index stream (trainClass.shuffled_index_stream + batched) decides WHICH samples form each global batch, identically
on every rank; this module moves a rank's shard of each batch to its GPU ahead of time:

  producer thread:  gather the shard's rows of X / HR / mask into a PINNED staging slot  ->  cudaMemcpyAsync on a private
                    copy stream into the slot's device buffers  ->  record a "ready" event
  consumer (the training loop):  the compute stream waits on "ready", runs the step on the slot's device tensors, and
                    records a "done" event when it asks for the next batch; the producer's next copy into that slot waits
                    on "done" (stream-level, no host sync).

`depth` slots (default 3) keep one batch in flight on the copy engine while another is being computed on, which is what
prefetch(AUTOTUNE) buys the reference.  With device=None the same code runs on host arrays without CUDA (CPU tests)."""
from __future__ import annotations

import queue
import threading
from typing import Iterable, Iterator, Sequence, Tuple

import numpy as np

from . import parallel

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class PrefetchLoader:
    def __init__(self, arrays: Sequence[np.ndarray], index_batches: Iterable[np.ndarray], device=None, depth: int = 3,
                 rank: int = None, world_size: int = None, max_batch: int = None):
        """arrays: host arrays with a common leading sample axis (X [N,S,S,T,1] f32, HR [N,..] f32, mask bool/uint8).
        index_batches: iterable of global-batch index arrays.  Yields (global_batch_size, (dev_or_host arrays of the shard))."""
        self.arrays = [np.asarray(a) for a in arrays]
        self.arrays = [a.view(np.uint8) if a.dtype == np.bool_ else a for a in self.arrays]
        self.batches = index_batches
        self.device = device
        self.depth = max(2, int(depth))
        r, w = parallel.world()
        self.rank = r if rank is None else rank
        self.ws = w if world_size is None else world_size
        self.max_batch = max_batch
        self._q: "queue.Queue" = queue.Queue(maxsize=self.depth - 1)
        self._free: "queue.Queue" = queue.Queue()
        self._stop = threading.Event()
        self._thread = None
        self._slots = None

    # ---- slots
    def _alloc(self, cap: int):
        self._slots = []
        for _ in range(self.depth):
            if self.device is None:
                host = [np.empty((cap,) + a.shape[1:], a.dtype) for a in self.arrays]
                self._slots.append({"host": host, "dev": host, "ready": None, "done": None})
            else:
                pinned = [torch.empty((cap,) + a.shape[1:], dtype=torch.from_numpy(a[:0]).dtype).pin_memory() for a in self.arrays]
                dev = [torch.empty_like(p, device=self.device) for p in pinned]
                self._slots.append({"host": [p.numpy() for p in pinned], "pinned": pinned, "dev": dev,
                                    "ready": torch.cuda.Event(), "done": None})
        for k in range(self.depth):
            self._free.put(k)

    def _shard(self, idx: np.ndarray) -> np.ndarray:
        lo, hi = parallel.shard_bounds(len(idx), self.rank, self.ws)
        return np.sort(idx[lo:hi]) if hi > lo else idx[:0]

    def _produce(self):
        copy_stream = torch.cuda.Stream(self.device) if self.device is not None else None
        try:
            for idx in self.batches:
                if self._stop.is_set():
                    break
                idx = np.asarray(idx)
                sel = self._shard(idx)
                n = len(sel)
                if self._slots is None:
                    lo, hi = parallel.shard_bounds(self.max_batch or len(idx), 0, self.ws)
                    self._alloc(max(hi - lo, n, 1))
                k = self._free.get()
                if k is None:
                    break
                s = self._slots[k]
                if copy_stream is not None and s["done"] is not None:
                    s["ready"].synchronize()                       # the slot's previous H2D copy has left the pinned buffer
                for a, h in zip(self.arrays, s["host"]):
                    np.take(a, sel, axis=0, out=h[:n])
                if copy_stream is not None:
                    with torch.cuda.stream(copy_stream):
                        if s["done"] is not None:
                            copy_stream.wait_event(s["done"])      # the step that last used this slot has been issued and must finish first
                        for p, d in zip(s["pinned"], s["dev"]):
                            d[:n].copy_(p[:n], non_blocking=True)
                        s["ready"].record(copy_stream)
                self._q.put((k, n, len(idx)))
        except BaseException as e:          # surfaced in the consumer
            self._q.put(e)
            return
        self._q.put(None)

    def __iter__(self) -> Iterator[Tuple[int, tuple]]:
        self._thread = threading.Thread(target=self._produce, name="pv-prefetch", daemon=True)
        self._thread.start()
        prev = None
        try:
            while True:
                item = self._q.get()
                if prev is not None:        # the caller has issued its work on the previous slot: hand it back
                    s = self._slots[prev]
                    if self.device is not None:
                        s["done"] = torch.cuda.Event()
                        s["done"].record(torch.cuda.current_stream(self.device))
                    self._free.put(prev)
                    prev = None
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                k, n, gb = item
                s = self._slots[k]
                if self.device is not None:
                    torch.cuda.current_stream(self.device).wait_event(s["ready"])
                prev = k
                yield gb, tuple(d[:n] for d in s["dev"])
        finally:
            self.close()

    def close(self):
        self._stop.set()
        self._free.put(None)
        while self._thread is not None and self._thread.is_alive():
            try:
                self._q.get(timeout=0.05)          # unblock a producer waiting on a full queue
            except queue.Empty:
                pass
            self._thread.join(timeout=0.05)
        self._thread = None
