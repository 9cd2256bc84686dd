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


class PeerAvgExchange:
    """``reduce_mean`` of the (2,) avg-factor tensor as one kernel over NVLink peer memory
    (``erd_avg_exchange``): every rank owns a small symmetric buffer all peers have mapped.
    Plumbing only: torch symmetric memory allocates and maps the buffers.  ``create`` returns
    None (the caller then keeps the NCCL all-reduce) when the process group is not initialised,
    has one rank, spans more than one node, or symmetric memory cannot be set up.  A peer that
    never arrives makes both factors NaN (every loss of the step is then NaN) and sets a status word."""

    def __init__(self, lib, buf, handle, rank: int, world_size: int):
        import ctypes as C
        self.lib, self.buf, self.handle = lib, buf, handle
        self.rank, self.world_size = rank, world_size
        self.peers = (C.c_void_p * world_size)(*[int(p) for p in handle.buffer_ptrs])

    @classmethod
    def create(cls, lib, device: torch.device):
        import os
        rank, ws = world()
        if ws < 2 or ws > 64 or os.environ.get('ERD_PEER_EXCHANGE', '1') == '0':
            return None
        if torch.device(device).type != 'cuda' or 'nccl' not in str(dist.get_backend()):
            return None
        # peer memory is a single-node mechanism.  Launchers that do not export LOCAL_WORLD_SIZE
        # (mmengine's slurm launcher) are asked directly: do all ranks report the same host?
        if 'LOCAL_WORLD_SIZE' in os.environ:
            if int(os.environ['LOCAL_WORLD_SIZE']) != ws:
                return None
        else:
            import socket
            hosts = [None] * ws
            dist.all_gather_object(hosts, socket.gethostname())
            if len(set(hosts)) != 1:
                return None
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = int(lib.erd_avg_exchange_bytes())
            with torch.cuda.device(device):
                buf = symm.empty((nbytes + 3) // 4, dtype=torch.int32, device=device)
                buf.zero_()
                handle = symm.rendezvous(buf, dist.group.WORLD)
                torch.cuda.synchronize(device)
            dist.barrier()                   # every buffer is zero before anyone stores into it
            return cls(lib, buf, handle, rank, ws)
        except Exception as e:               # noqa: BLE001 -- any failure means "use NCCL"
            import warnings
            warnings.warn(f'erd_b200: peer-memory avg exchange unavailable ({e!r}); using NCCL all_reduce')
            return None

    def reduce_mean_(self, t: torch.Tensor) -> torch.Tensor:
        assert t.dtype == torch.float32 and t.numel() == 2 and t.is_contiguous()
        rc = self.lib.erd_avg_exchange(t.data_ptr(), self.peers, self.rank, self.world_size,
                                       torch.cuda.current_stream(t.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f'erd_avg_exchange failed ({rc})')
        return t

    def timed_out(self) -> bool:
        """True if a peer never arrived in some earlier exchange (host sync; for tests / teardown)."""
        return bool(self.buf[int(self.lib.erd_avg_exchange_bytes()) // 4 - 3].item())


_peer_exchange = None
_peer_exchange_tried = False


def peer_exchange(lib, device: torch.device):
    """The process-wide PeerAvgExchange (created on first use -- a collective call, made by every
    rank at its first reduction, outside any CUDA-graph capture) or None."""
    global _peer_exchange, _peer_exchange_tried
    if not _peer_exchange_tried:
        _peer_exchange_tried = True
        _peer_exchange = PeerAvgExchange.create(lib, device)
        if world()[1] > 1:
            # all or nothing: a rank that could not set it up must not leave the others spinning
            ok = torch.tensor([1 if _peer_exchange is not None else 0], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                _peer_exchange = None
    return _peer_exchange


def shard_images(num_global: int, rank: int, world_size: int) -> List[int]:
    """Contiguous block of image indices owned by ``rank`` (DDP shards the batch by
    image; the path needs no halo and no data-path collective)."""
    base, rem = divmod(num_global, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))
