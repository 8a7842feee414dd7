"""The N>1 path on CPU: two gloo ranks shard a frame by sample range, each fills an int64 accumulator, one
reduce(SUM) gives exactly the single-rank accumulator.  (The rendering itself needs a GPU; here each rank's
"render" is a deterministic integer function of (pixel, sample), which is all the reduce logic can see.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cs397raytracingsp22_b200 import distributed as D


def _fake_shard(w, h, s0, s1):
    """Stand-in for rt_render_accum: integer contribution of samples [s0, s1) to every pixel channel."""
    p = torch.arange(w * h * 4, dtype=torch.int64)
    acc = torch.zeros(w * h * 4, dtype=torch.int64)
    for s in range(s0, s1):
        acc += ((p * 2654435761 + s * 40503) % 1000003) - 500000
    return acc


def _worker(rank, world, port, w, h, spp, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s0, s1 = D.sample_range(rank, world, 0, spp)
        acc = _fake_shard(w, h, s0, s1)
        D.reduce_accum(acc, dst=0)
        if rank == 0:
            q.put(acc.numpy().copy())
        acc2 = _fake_shard(w, h, s0, s1)
        D.reduce_accum(acc2, dst=None)         # all_reduce variant: every rank ends with the frame
        q.put((rank, int(acc2.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_ranks_reduce_to_the_single_rank_accumulator():
    w, h, spp, world = 16, 9, 10, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, spp, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=90) for _ in range(world + 1)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    frame = next(g for g in got if isinstance(g, np.ndarray))
    want = _fake_shard(w, h, 0, spp).numpy()
    assert np.array_equal(frame, want)
    sums = dict(g for g in got if isinstance(g, tuple))
    assert sums[0] == sums[1] == int(want.sum())


def test_reduce_is_a_no_op_without_a_process_group():
    acc = torch.arange(8, dtype=torch.int64)
    assert D.reduce_accum(acc) is acc



@pytest.mark.parametrize("world,w,h,tile", [(2, 50, 37, 16), (3, 33, 20, 8), (8, 1920, 1080, 16)])
def test_tile_shards_partition_the_frame(world, w, h, tile):
    """Every tile (ragged ones at the right and bottom edges included) has exactly one owner, and no rank owns more
    than its share plus one diagonal's worth - the host-side mirror of plan_shard() in csrc/rt_api.cu."""
    tx, ty = (w + tile - 1) // tile, (h + tile - 1) // tile
    seen = {}
    for r in range(world):
        for t in D.tiles_of_rank(r, world, w, h, tile):
            assert t not in seen
            seen[t] = r
    assert len(seen) == tx * ty
    counts = [sum(1 for v in seen.values() if v == r) for r in range(world)]
    assert max(counts) - min(counts) <= max(tx, ty) // world + 1
    # neighbours in a row and in a column belong to different ranks (diagonals, not columns or rows)
    assert all(seen[(x, y)] != seen[(x + 1, y)] for y in range(ty) for x in range(tx - 1))
    assert all(seen[(x, y)] != seen[(x, y + 1)] for y in range(ty - 1) for x in range(tx))
