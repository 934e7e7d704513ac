"""Pose sharding across the GPUs of one box: contiguous blocks of the pose index range,
costmap and footprints replicated on every rank, no collective on the data path; one final gather of
the per-rank result slices (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def partition(count: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced block of [0, count) owned by `rank`: returns (first, n).  The first
    `count % world_size` ranks get one extra pose; blocks are disjoint and cover the range."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, extra = divmod(count, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def max_block(count: int, world_size: int) -> int:
    return (count + world_size - 1) // world_size


def gather_slices(local, count: int, group=None):
    """All-gathers the ranks' result slices (1-D tensors produced for `partition(count, world, rank)`) into
    the full `count`-long tensor, in pose order, on every rank.  Works on CPU (gloo) and CUDA (nccl) tensors."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    first, n = partition(count, world, rank)
    if local.numel() != n:
        raise ValueError(f"rank {rank} holds {local.numel()} results, expected {n}")
    pad = max_block(count, world)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[:n] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = torch.empty(count, dtype=local.dtype, device=local.device)
    for r in range(world):
        f, m = partition(count, world, r)
        out[f:f + m] = parts[r][:m]
    return out
