"""Partitioning across the GPUs of one box (one process per GPU, torch.distributed).

Two levels; the reference has neither (nn.DataParallel only, D6):
  * slides are independent units -> data parallel; the only collective is the gradient all-reduce
    (``step.allreduce_gradients``; NCCL over NVLink in production, gloo in the CPU tests);
  * a giant bag is split into contiguous patch shards, one per rank; every pooling block exchanges
    its (pooled, lse) partial state once (an all-gather of P x 257 floats per bag) and merges it with
    the log-sum-exp kernel; the backward needs one all-reduce of dq~ per block plus the usual
    gradient all-reduce for dW1/db1 (the token algebra is replicated on every rank).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

TILE = 64      # patch rows per pooling tile (csrc/pool.cu): shard boundaries are kept tile-aligned


def shard_bounds(n: int, world: int, align: int = TILE) -> List[Tuple[int, int]]:
    """Contiguous, tile-aligned row ranges covering [0,n); trailing shards may be empty."""
    tiles = (n + align - 1) // align
    per = (tiles + world - 1) // world
    out = []
    for r in range(world):
        a = min(n, r * per * align)
        b = min(n, (r + 1) * per * align)
        out.append((a, b))
    return out


def assign_slides(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-first balancing of slides over ranks (bags vary from 1 to 10k+ patches)."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(v) for v in out]


def gather_partials(pooled: torch.Tensor, lse: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the per-rank (pooled (B,P,D), lse (B,P)) -> (B,W,P,D), (B,W,P) in rank order."""
    world = dist.get_world_size(group)
    packed = torch.cat([pooled, lse.unsqueeze(-1)], dim=-1).contiguous()          # (B,P,D+1): one message
    bufs = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(bufs, packed, group=group)
    allp = torch.stack(bufs, dim=1)                                                # (B,W,P,D+1)
    return allp[..., :-1].contiguous(), allp[..., -1].contiguous()


def merge_shards(pooled: torch.Tensor, lse: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-rank partial pooling state -> whole-bag (pooled, lse), identical on every rank."""
    from . import kernels
    part_p, part_l = gather_partials(pooled, lse, group)
    return kernels.lse_merge(part_p, part_l)


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    dist.all_reduce(t, group=group)
    return t
