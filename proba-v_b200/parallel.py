"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

Semantics follow the reference's (unfinished) MirroredStrategy draft, debug/trainClassMultiGPU0.py:67-84,137-140,
162-178: per-replica batch = cfg batch_size, global batch = batch_size x replicas, loss scaled by 1/global batch,
gradients summed across replicas, metrics averaged.  The exchange step is ONE all-reduce of the flat gradient arena
(plus 2 metric scalars appended), never per-variable.
"""
from __future__ import annotations

import os
from typing import Tuple

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Initialise from torchrun's RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* env; returns (rank, world, local_rank)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": torch.device(f"cuda:{local}")} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=ws, **kw)
    return rank, ws, local


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, as-equal-as-possible shard [lo, hi) of n samples for `rank` (first n % world get one extra)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def grad_scale(global_batch: int) -> float:
    """Each rank back-propagates sum_b loss_b / global_batch, so a SUM all-reduce yields the global-batch mean
    even when shards are unequal (the reference's partial last batch, utils/utils.py:32-34)."""
    return 1.0 / float(global_batch)


def allreduce_sum_(t):
    """In-place SUM all-reduce of a flat tensor (the gradient arena); no-op for world 1."""
    _, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_sum_async_(t):
    """Asynchronous in-place SUM all-reduce of one gradient bucket; returns the work handle (None for world 1).  With
    NCCL the collective is enqueued on NCCL's own stream behind the producer kernels of the current stream, so compute
    launched afterwards overlaps it; wait() makes the current stream wait for it."""
    _, ws = world()
    if ws > 1:
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
    return None


def broadcast_(t, src: int = 0):
    _, ws = world()
    if ws > 1:
        dist.broadcast(t, src=src)
    return t


def reduce_metrics(loss_mean_local: float, cpsnr_mean_local: float, n_local: int, device=None) -> Tuple[float, float]:
    """Global-batch means of (loss, cPSNR) from per-rank means and counts (strategy.reduce(MEAN) in the draft)."""
    _, ws = world()
    if ws == 1:
        return loss_mean_local, cpsnr_mean_local
    v = torch.tensor([loss_mean_local * n_local, cpsnr_mean_local * n_local, float(n_local)], dtype=torch.float64, device=device)
    dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return float(v[0] / v[2]), float(v[1] / v[2])
