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
