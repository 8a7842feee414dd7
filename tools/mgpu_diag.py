"""Where does a strong-scaled frame spend its time?  (development only; run under torch.distributed.run)

Per rank and per shard mode: CUDA-event times of render / collective / resolve inside back-to-back frames of the C4
workload, then the collectives alone on an accumulator-sized buffer: reduce, all_reduce, and the exchange of owned
pixels (gather of 1/N of the buffer) that a tile shard could use instead of a reduce.
usage: python -m torch.distributed.run --nproc-per-node N tools/mgpu_diag.py [--workload c4] [--frames 10]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import torch
import torch.distributed as dist

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--frames", type=int, default=10)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sys.argv = ["bench.py", "--gpus", str(world), "--workload", a.workload]
    args = bench.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    for mode in ("tiles", "samples"):
        job = bench.Job(args, None, rank, world, local, dev, shard=mode)
        D = job.D
        for _ in range(3):
            job.step(job.opts())
        dist.barrier(); torch.cuda.synchronize()
        marks = []
        t_all0, t_all1 = ev(), ev()
        t_all0.record()
        for _ in range(a.frames):
            e = [ev() for _ in range(4)]
            flush.zero_()
            job.accum.zero_()
            e[0].record()
            D.render_shard(job.g, job.cam, job.opts(), job.accum)
            e[1].record()
            D.reduce_accum(job.accum, dst=0)
            e[2].record()
            if rank == 0:
                D.resolve(job.g, job.cam, job.accum, job.total_spp)
            e[3].record()
            marks.append(e)
        t_all1.record()
        dist.barrier(); torch.cuda.synchronize()
        n = len(marks)
        ren = sum(m[0].elapsed_time(m[1]) for m in marks) / n
        red = sum(m[1].elapsed_time(m[2]) for m in marks) / n
        res = sum(m[2].elapsed_time(m[3]) for m in marks) / n
        tot = t_all0.elapsed_time(t_all1) / n
        t = torch.tensor([ren, red, res, tot], dtype=torch.float64, device=dev)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        if rank == 0:
            print(f"== {a.workload} {mode}: per rank  render / collective (incl. waiting for the slowest rank) / resolve / frame  [ms]")
            for r, x in enumerate(g):
                print(f"   rank {r}: " + "  ".join(f"{v:8.3f}" for v in x.tolist()))
        nbytes = job.accum.numel() * 8
        job.close()

    # the collectives alone
    n64 = nbytes // 8
    buf = torch.ones(n64, dtype=torch.int64, device=dev)
    part = n64 // world
    mine = torch.ones(part, dtype=torch.int64, device=dev)
    recv = [torch.empty(part, dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None
    full = torch.empty(n64, dtype=torch.int64, device=dev)

    def timeit(name, fn, reps=10):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        a0, a1 = ev(), ev()
        a0.record()
        for _ in range(reps):
            fn()
        a1.record()
        torch.cuda.synchronize()
        ms = a0.elapsed_time(a1) / reps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"   {name:58s} {float(t.item()):8.3f} ms  ({nbytes / 1e6:.0f} MB buffer)")

    if rank == 0:
        print("== collectives alone, back to back, max over ranks")
    timeit("reduce(SUM, int64) to rank 0", lambda: dist.reduce(buf, dst=0))
    timeit("all_reduce(SUM, int64)", lambda: dist.all_reduce(buf))
    timeit("reduce(SUM) of the same bytes viewed as int32", lambda: dist.reduce(buf.view(torch.int32), dst=0))
    timeit("gather of 1/N of the buffer to rank 0 (owned pixels)", lambda: dist.gather(mine, recv, dst=0))
    timeit("reduce_scatter(SUM, int64) + gather of the parts", lambda: (dist.reduce_scatter_tensor(mine, buf), dist.gather(mine, recv, dst=0)))
    timeit("all_gather of 1/N parts into a full buffer", lambda: dist.all_gather_into_tensor(full[:part * world], mine))
    timeit("barrier", lambda: dist.barrier())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
