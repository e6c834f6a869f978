"""Multi-GPU host plumbing (one process per GPU): how work is partitioned and how the NCCL communicator of a
ctx is bootstrapped.  The reference has no distributed code; this is additive and only
covers what shards naturally:

  * row blocks of one tall matrix (BASELINE config 4): `row_shard` + `init_comm` + `Context.set_row_shard`, then the
    ordinary fused front-ends on the local block -- the library all-reduces the sketch once per round;
  * batches of independent blocks (BASELINE config 5): `block_shard`, no collective at all.

torch.distributed is used for exactly one thing: shipping rank 0's 128-byte NCCL unique id to the other ranks.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

from ._binding import Context, lib


def row_shard(m: int, rank: int, world: int) -> Tuple[int, int]:
    """(row0, m_local) of this rank's contiguous row block; blocks are even-sized so every shard keeps the
    16-byte alignment the TMA path wants (the last blocks may be shorter or empty)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank/world")
    chunk = -(-m // world)
    chunk += chunk & 1
    row0 = min(rank * chunk, m)
    return row0, min(chunk, m - row0)


def block_shard(nblocks: int, rank: int, world: int) -> Tuple[int, int]:
    """(first block, count): contiguous groups of independent blocks, sizes differing by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank/world")
    base, rem = divmod(nblocks, world)
    b0 = rank * base + min(rank, rem)
    return b0, base + (1 if rank < rem else 0)


def broadcast_bytes(buf: bytes, src: int = 0, group=None, device=None) -> bytes:
    """Every rank returns rank `src`'s bytes (torch.distributed broadcast of a uint8 tensor; `device` must be a CUDA
    device under the nccl backend, None/cpu under gloo)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(buf), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def unique_id() -> bytes:
    buf = (C.c_ubyte * 128)()
    rc = lib.bra_comm_unique_id(buf)
    if rc != 0:
        raise RuntimeError(f"bra_comm_unique_id failed with status {rc} (is libnccl.so.2 loadable?)")
    return bytes(buf)


def init_comm(ctx: Context, rank: int, world: int, group=None, device=None) -> None:
    """Builds ctx's NCCL communicator: rank 0 draws the unique id, torch.distributed ships it."""
    uid = unique_id() if rank == 0 else bytes(128)
    if world > 1:
        uid = broadcast_bytes(uid, 0, group, device)
    buf = (C.c_ubyte * 128).from_buffer_copy(uid)
    ctx.check(lib.bra_comm_init(ctx.handle, buf, rank, world))
