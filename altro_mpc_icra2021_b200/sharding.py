"""Multi-GPU sharding of an instance batch.

MPC instances are independent, so the batch is cut into contiguous ranges, one per rank (one process
per GPU), and the solve path has no collective.  The only exchange is the final gather of the small
per-instance statistics (iterations, status, cost, c_max), done with torch.distributed: NCCL over
NVLink on GPUs, gloo in the CPU tests.  The reference has no counterpart (it solves one problem at a
time, single-threaded): this is the batched replacement for looping over problems.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [i0, i1) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    i0 = rank * base + min(rank, rem)
    return i0, i0 + base + (1 if rank < rem else 0)


def init_distributed(backend: Optional[str] = None):
    """Joins the torchrun rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT). Returns (rank, world)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def gather_stats(stats: Dict[str, np.ndarray], device: Optional[str] = None) -> Dict[str, np.ndarray]:
    """all_gather of per-instance statistics; every rank gets the concatenation in rank order.
    Shards may have different lengths (padded to the longest for the collective)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: np.asarray(v) for k, v in stats.items()}
    world = dist.get_world_size()
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    keys = sorted(stats)
    n_local = len(stats[keys[0]])
    lens = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(lens, torch.tensor([n_local], dtype=torch.int64, device=device))
    lens = [int(x.item()) for x in lens]
    nmax = max(lens)
    packed = torch.zeros((len(keys), nmax), dtype=torch.float64, device=device)
    for j, k in enumerate(keys):
        packed[j, :n_local] = torch.as_tensor(np.asarray(stats[k], dtype=np.float64), device=device)
    parts = [torch.zeros_like(packed) for _ in range(world)]
    dist.all_gather(parts, packed)
    out = {}
    for j, k in enumerate(keys):
        cat = np.concatenate([parts[r][j, :lens[r]].cpu().numpy() for r in range(world)])
        out[k] = cat.astype(np.asarray(stats[k]).dtype)
    return out


def max_over_ranks(value: float, device: Optional[str] = None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
