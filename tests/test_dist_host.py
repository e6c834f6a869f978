"""Host-side logic of the multi-GPU path on CPU: partitioning helpers, and the unique-id broadcast over a
world_size-2 gloo group (no GPU, no NCCL: the id is a fake 128-byte pattern)."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_shard_partitions_rows():
    import brapprox
    for m in (0, 1, 7, 4096, 1048576, 6001):
        for world in (1, 2, 3, 8):
            seen = 0
            for g in range(world):
                row0, ml = brapprox.row_shard(m, g, world)
                assert row0 == min(seen, m) or ml == 0
                assert row0 % 2 == 0 or ml == 0
                assert ml >= 0
                seen = max(seen, row0 + ml)
            assert seen == m
    with pytest.raises(ValueError):
        brapprox.row_shard(10, 2, 2)


def test_block_shard_balanced():
    import brapprox
    for nb in (0, 5, 16384, 1000):
        for world in (1, 2, 4, 8, 3):
            counts = [brapprox.block_shard(nb, g, world) for g in range(world)]
            assert sum(c for _, c in counts) == nb
            assert max(c for _, c in counts) - min(c for _, c in counts) <= 1
            nxt = 0
            for b0, c in counts:
                assert b0 == nxt
                nxt += c


def _bcast_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
    import torch.distributed as dist
    from brapprox._dist import broadcast_bytes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = bytes(range(128)) if rank == 0 else bytes(128)
    got = broadcast_bytes(uid, 0)
    q.put((rank, got == bytes(range(128))))
    dist.destroy_process_group()


def test_unique_id_broadcast_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, 29577, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
