"""Multi-GPU plumbing for the batch shard: one process per GPU, every rank runs the
whole hot path on its own OD samples (the B*N*N cells of different samples are independent given G); the only
exchange step is the mean all-reduce of the parameter gradients (< 200 KB) over NCCL / NVLink.

The reference has no distributed code at all; this is the B200-native addition.
"""
from __future__ import annotations

import os
from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl", device=None) -> Tuple[int, int]:
    """Join the torchrun rendezvous (RANK / WORLD_SIZE / MASTER_* from the environment). Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_items` owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_gradients(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """Average .grad over all ranks with ONE collective on a flat buffer. Parameters without a gradient
    contribute zeros (every rank must issue the same collective). Returns the number of reduced elements."""
    params = list(params)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or not params:
        return 0
    world = dist.get_world_size(group)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, group=group)
    flat /= world
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off
