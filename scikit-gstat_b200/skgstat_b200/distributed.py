"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the GPU
box, gloo in the CPU tests).  The pair triangle shards by tile range, kriging
targets by contiguous slices; the only data-path collectives are

  * ONE all-reduce(sum) of the per-lag partials [count | sum dz^2 | sum sqrt|dz|]
    after the fused pass (counts travel as exact integers in fp64, < 2^53),
  * all-reduce(sum) of bucket histograms / all-reduce(max) of the max key and an
    all-gather of the few candidates in the order-statistic passes.

Kriging needs no collective (outputs are concatenated by an all-gather of equal
slices only when the caller asks for the full field on every rank).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


LOCAL = "local"   # pseudo group: this process alone, whatever torch.distributed's state


def world(group=None) -> Tuple[int, int]:
    """(rank, world_size) of the process group; (0, 1) when not initialised or for LOCAL."""
    if isinstance(group, str) and group == LOCAL:
        return 0, 1
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) share of `total` work units for `rank`."""
    base, rem = divmod(int(total), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_max_(t: torch.Tensor, group=None) -> torch.Tensor:
    if world(group)[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def allreduce_max_int(value: int, device, group=None) -> int:
    """max over the ranks of a host integer (error decisions must be collective)."""
    if world(group)[1] == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def allgather_varlen(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate 1-D tensors of different lengths from all ranks (rank order)."""
    rank, ws = world(group)
    if ws == 1:
        return t
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros(cap, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    bufs = [torch.zeros_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])
