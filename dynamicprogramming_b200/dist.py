"""
State-range sharding across the GPUs of one box (new; the reference is single-GPU).

One process per GPU (torchrun).  torch.distributed is used only for plumbing:
rank/world discovery and shipping the ncclUniqueId from rank 0 to the other
ranks; the data path (per-sweep V exchange, residual max, changed-count sum)
runs inside libdpb200.so on its own NCCL communicator, enqueued on the engine's
stream and captured into the sweep graphs.

Partition: rank r owns the contiguous flat-index range [r*N//W, (r+1)*N//W)
(dim 0 slowest => slabs of dim 0).  Exchange: after each sweep a rank receives,
from every peer, only the sub-range of that peer's slice that its transition
rows reference (`plan_exchange`); an all-gather is the degenerate case.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi


def shard_range(rank: int, world: int, n_states: int) -> tuple[int, int]:
    """Flat-index range owned by `rank` (same arithmetic as pi_create)."""
    return rank * n_states // world, (rank + 1) * n_states // world


def owner_of(index: int, world: int, n_states: int) -> int:
    """Rank whose range contains flat index `index`."""
    return ((index + 1) * world - 1) // n_states


def plan_exchange(base: np.ndarray, span: int, rank: int, world: int, n_states: int) -> list[tuple[int, int]]:
    """Host mirror of the engine's need-range computation.

    `base` holds the (non-negative) base indices of the rows this rank owns, for
    all actions; every row touches flat indices [base, base + span] where span =
    sum of strides.  Returns, per peer, the half-open global range [lo, hi) this
    rank must receive from it ((0, 0) if nothing)."""
    need = [(0, 0)] * world
    base = np.asarray(base, dtype=np.int64)
    base = base[base >= 0]
    if base.size == 0:
        return need
    for r in range(world):
        if r == rank:
            continue
        rlo, rhi = shard_range(r, world, n_states)
        touch = (base + span >= rlo) & (base < rhi)
        if not touch.any():
            continue
        b = base[touch]
        lo = max(int(b.min()), rlo)
        hi = min(int(b.max()) + span + 1, rhi)
        if hi > lo:
            need[r] = (lo, hi)
    return need


def init_process_group(backend: str | None = None):
    """torch.distributed init from the torchrun environment (idempotent)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
    return dist


def make_shard(device: int | None = None) -> tuple[int, int, bytes] | None:
    """(rank, world, ncclUniqueId bytes) for the engines' `shard=` argument, or
    None for a single process.  Rank 0 creates the id; it is broadcast through
    the torch process group."""
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    dist = init_process_group()
    rank = dist.get_rank()
    buf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _ffi.check(_ffi.lib().pi_nccl_unique_id(_ffi.ptr(buf)))
    t = torch.from_numpy(buf)
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)) if device is None else device)
        t = t.to(dev)
        dist.broadcast(t, src=0)
        buf = t.cpu().numpy()
    else:
        dist.broadcast(t, src=0)
    return rank, world, bytes(buf.tobytes())
