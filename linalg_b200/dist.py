"""One-process-per-GPU plumbing for the sharded paths.

The data path never goes through this module: batched problems are split by batch index with no
communication, and the two row-sharded paths (TSQR, Gram/SVD) exchange their 128 x 128 factors
inside the C library with NCCL (``lq_comm_*``).  What lives here is the host-side control plane a
launcher such as ``torchrun`` needs: rank discovery from the environment, the hand-off of the NCCL
unique id from rank 0 to the other ranks, a barrier and a max-over-ranks reduction for timings.
``torch.distributed`` (backend ``gloo``: host tensors only) is used for that and imported lazily,
so single-GPU users never import torch.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .utils import shard_bounds


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int


def rank_info() -> RankInfo:
    env = os.environ
    return RankInfo(int(env.get("RANK", "0")), int(env.get("WORLD_SIZE", "1")), int(env.get("LOCAL_RANK", env.get("RANK", "0"))))


def init_control_plane(backend: str = "gloo") -> RankInfo:
    """Join the launcher's rendezvous (MASTER_ADDR/MASTER_PORT) when WORLD_SIZE > 1."""
    info = rank_info()
    if info.world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            dist.init_process_group(backend=backend, rank=info.rank, world_size=info.world)
    return info


def shutdown_control_plane():
    info = rank_info()
    if info.world > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


def barrier():
    if rank_info().world > 1:
        import torch.distributed as dist

        dist.barrier()


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Every rank returns rank ``src``'s ``payload`` (exactly ``nbytes`` long)."""
    info = rank_info()
    if info.world == 1:
        assert payload is not None
        return bytes(payload)
    import torch
    import torch.distributed as dist

    buf = torch.zeros(nbytes, dtype=torch.uint8)
    if info.rank == src:
        assert payload is not None and len(payload) == nbytes
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.numpy().tobytes())


def max_over_ranks(value: float) -> float:
    if rank_info().world == 1:
        return float(value)
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    if rank_info().world == 1:
        return float(value)
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allgather_object(obj):
    """Every rank returns the list of all ranks' (picklable, small) objects, in rank order."""
    info = rank_info()
    if info.world == 1:
        return [obj]
    import torch.distributed as dist

    out = [None] * info.world
    dist.all_gather_object(out, obj)
    return out


def init_comm(ctx, info: RankInfo | None = None):
    """Create the NCCL communicator of ``ctx`` (collective over all ranks of the control plane)."""
    info = info or rank_info()
    if info.world == 1:
        ctx.call("lq_comm_init", 1, 0, None)
        return
    from ._native import check

    ident = None
    if info.rank == 0:
        raw = (C.c_char * 128)()
        check(ctx.lib, None, ctx.lib.lq_comm_unique_id(C.cast(raw, C.c_void_p)), "lq_comm_unique_id")
        ident = bytes(raw.raw)
    ident = broadcast_bytes(ident, 128, src=0)
    raw = (C.c_char * 128).from_buffer_copy(ident)
    ctx.call("lq_comm_init", info.world, info.rank, C.cast(raw, C.c_void_p))


def my_batch_slice(total: int, info: RankInfo | None = None):
    """Batch-sharded paths (cfg2 / cfg3): contiguous split of the batch index, no communication."""
    info = info or rank_info()
    return shard_bounds(total, info.world, info.rank)


def my_row_slice(rows: int, info: RankInfo | None = None, align: int = 1):
    """Row-sharded paths (cfg5): contiguous row blocks; the R / Gram exchange happens in the library."""
    info = info or rank_info()
    return shard_bounds(rows, info.world, info.rank, align)


__all__ = [
    "RankInfo", "rank_info", "init_control_plane", "shutdown_control_plane", "barrier", "broadcast_bytes",
    "max_over_ranks", "sum_over_ranks", "allgather_object", "init_comm", "my_batch_slice", "my_row_slice",
]
