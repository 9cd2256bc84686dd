"""Data-parallel plumbing of the path: the scalar mean all-reduce and image sharding.

Reference: ``reduce_mean`` (mmdet/utils/dist_utils.py:59-65), called twice per step at
gfl_head_increment_erd.py:390-391 and :406-407; here both operands travel in one
2-float tensor that stays on the device (no ``.item()``).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def reduce_mean_(t: torch.Tensor) -> torch.Tensor:
    """In-place mean over ranks: ``t.div_(world).all_reduce(SUM)``; passthrough when the
    process group is not initialised (dist_utils.py:61-62)."""
    _, ws = world()
    if ws > 1:
        t.div_(ws)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_images(num_global: int, rank: int, world_size: int) -> List[int]:
    """Contiguous block of image indices owned by ``rank`` (DDP shards the batch by
    image; the path needs no halo and no data-path collective)."""
    base, rem = divmod(num_global, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))
