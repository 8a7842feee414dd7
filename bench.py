#!/usr/bin/env python
"""bench.py — headline benchmark: Msamples/s (and Mrays/s) of the path-tracing hot path on B200.

  python bench.py --gpus 1 --steps K --warmup W            one GPU
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   one rank per GPU
  python bench.py --impl reference ...                     the reference-equivalent CPU renderer (oracle port)

A "step" is one full render of the workload (default: BASELINE config C4, the textured drone scene of the
reference's run(), 1920x1080, 1024 spp, depth 10).  With N ranks the scaling is STRONG: the same one frame is cut
into interleaved 16x16 tiles (rank r renders tiles r, r+N, ...), the exact int64 accumulators are summed with one NCCL
reduce and rank 0 resolves the image - bit-identical to the single-GPU frame.  Prints ONE JSON line (rank 0).  The
`configs` block of that line carries the other BASELINE configurations at their full sizes (C1-C3 and C5 at N=1; C5,
tile- and sample-sharded, at N>1).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

SEED = 0x5EED
METRIC = "samples_per_second"
UNIT = "Msamples/s"
TILE = 16                      # strong-scaling tile edge: 8 160 tiles at 1080p, so 8 interleaved shards balance to a few %
B200_SMS, LANES_PER_SM_CLK = 148, 128   # 4 schedulers x 32 lanes: peak thread-instructions per SM per clock
ENGINE_NAMES = {1: "wavefront", 2: "megakernel"}   # rt_stats.engine (RT_ENGINE_*), summed over the steps of a timed region


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--shard", default="tiles", choices=["tiles", "samples", "weak"],
                    help="N>1: tiles / samples = strong scaling of one frame (default tiles); weak = every rank renders "
                         "the full spp with its own key")
    ap.add_argument("--tile", type=int, default=TILE)
    ap.add_argument("--engine", default="auto", choices=["auto", "wavefront", "megakernel"])
    ap.add_argument("--ray-sort", default="auto", choices=["auto", "off", "on"])
    ap.add_argument("--work-order", default="auto", choices=["auto", "pixel", "sample", "grouped"])
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--no-counters", action="store_true", help="skip the counted pass (profiling runs; zeroes the roofline)")
    ap.add_argument("--wavefront", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (invalidates the headline)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--emulate-shards", type=int, default=0, help="diagnostics: render one shard of this many (with --shard)")
    ap.add_argument("--emulate-rank", type=int, default=0, help="diagnostics: which shard --emulate-shards renders")
    ap.add_argument("--depth", type=int, default=0, help="override path_depth (diagnostics; invalidates the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other BASELINE configurations)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def scene_for(args, name=None):
    from cs397raytracingsp22_b200 import scenes
    kw = {}
    if name is None:                      # the main workload takes the diagnostic overrides; the configs block never does
        name = args.workload
        if args.spp:
            kw["spp"] = args.spp
        if args.width:
            kw["width"] = args.width
        if args.height:
            kw["height"] = args.height
        if args.depth:
            kw["depth"] = args.depth
    return scenes.make_scene(name, **kw), scenes.DESCRIPTIONS[name]


def parallelism(args) -> str:
    n = args.gpus
    if n == 1:
        return "single GPU"
    if args.shard == "weak":
        return f"{n} GPUs, every rank renders the full frame at the configured spp with its own Philox key; one NCCL int64 reduce per frame"
    how = f"interleaved {args.tile}x{args.tile} tiles" if args.shard == "tiles" else "contiguous sample ranges"
    return f"{n} GPUs, ONE frame cut into {how}; one NCCL int64 reduce per frame"


def workload_config(args, sc, desc) -> dict:
    """Declarative description of the job - identical for the GPU arm and the reference arm."""
    cam = sc.camera
    return {
        "workload": f"{args.workload}: {desc}",
        "width": cam.screen_width, "height": cam.screen_height, "spp": cam.aa_sample_count,
        "path_depth": cam.path_depth, "objects": len(sc.objects), "seed": SEED,
        "parallelism": parallelism(args),
        "total_spp": cam.aa_sample_count * (args.gpus if (args.gpus > 1 and args.shard == "weak") else 1),
        "l2": "flush (256 MiB memset between timed iterations)",
    }


def scaling_of(args) -> str:
    return "weak" if (args.gpus > 1 and args.shard == "weak") else "strong"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_render_sample(sc, n_spp: int, first: int = 0):
    """The reference-equivalent CPU renderer (oracle, reference-tree mode, OpenMP over rows like tracing.rs:228) on
    the full-resolution frame for sample indices [first, first+n_spp), on ALL host cores whatever OMP_NUM_THREADS says
    (torch.distributed.run sets it to 1).  Returns (samples, rays, seconds, threads used)."""
    import oracle_ffi as O
    b = O.lower_to_oracle(sc)
    cam = sc.camera.to_c()
    cores = host_cores()
    t0 = time.perf_counter()
    _, _, st = b.render(cam, seed=SEED, mode=O.MODE_REF_TREE, sample_begin=first, sample_end=first + n_spp,
                        nthreads=cores, want_linear=True, want_rgb8=False)
    dt = time.perf_counter() - t0
    used = O.load().orc_num_threads()
    b.close()
    return int(st.samples), int(st.rays), dt, used


def cpu_baseline(sc, target_seconds: float) -> dict:
    s, r, dt, cores = cpu_render_sample(sc, 1)
    n = max(1, min(int(target_seconds / max(dt, 1e-3)) - 1, sc.camera.aa_sample_count - 1, 64))
    if n >= 1 and dt < target_seconds * 0.6:
        s2, r2, dt2, _ = cpu_render_sample(sc, n, first=1)
        s, r, dt = s + s2, r + r2, dt + dt2
    cam = sc.camera
    return {"value": s / dt / 1e6, "unit": UNIT, "cores": cores, "host_cores": host_cores(), "kind": "port",
            "rays_per_sec_M": r / dt / 1e6, "seconds": dt,
            "sample": f"full {cam.screen_width}x{cam.screen_height} frame, {s // (cam.screen_width * cam.screen_height)} "
                      f"of {cam.aa_sample_count} sample indices per pixel, oracle in reference-tree mode, OpenMP "
                      f"schedule(dynamic) over rows",
            "note": "upper bound on the Rust reference's speed: no per-box-hit Arc allocation, no virtual dispatch "
                    "(BASELINE.md §3)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sc, desc = scene_for(args)
    for _ in range(args.warmup):
        cpu_render_sample(sc, 1)
    tot_s = tot_r = 0
    tot_t = 0.0
    cores = 1
    for k in range(args.steps):
        s, r, dt, cores = cpu_render_sample(sc, 1, first=k % sc.camera.aa_sample_count)
        tot_s += s; tot_r += r; tot_t += dt
    val = tot_s / tot_t / 1e6
    cam = sc.camera
    sample = (f"each step = full {cam.screen_width}x{cam.screen_height} frame at 1 of {cam.aa_sample_count} sample "
              f"indices per pixel ({cam.screen_width * cam.screen_height} paths), oracle port in reference-tree mode, "
              f"{cores} OpenMP threads on {host_cores()} host cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_t / max(args.steps, 1) * 1e3, "higher_is_better": True,
        "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sc, desc),
        "rays_per_sec_M": tot_r / tot_t / 1e6,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "host_cores": host_cores(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if cores < host_cores():
        line["cpu_baseline"]["void"] = f"only {cores} of {host_cores()} host cores were used: do not form a ratio with this line"
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------ GPU arm
def trace_algorithmic_bytes(st) -> float:
    """Algorithmic bytes of ALL closest-hit work of one frame: 32 B per BVH node fetched,
    48 B per triangle record, 96 B of transforms per instance entered, 32 B per analytic primitive record, plus the
    wavefront's own traffic in k_trace: 48 B ray read + 20 B hit record written per ray."""
    return (32.0 * st["nodes_visited"] + 48.0 * st["tris_tested"] + 96.0 * st["instances_entered"]
            + 32.0 * st["prims_tested"] + 68.0 * st["rays"])


class Job:
    """One workload on this rank's GPU: scene committed, accumulator allocated, shard options per pass."""

    def __init__(self, args, name, rank, world, local, dev, shard=None):
        import torch
        from cs397raytracingsp22_b200 import _ffi, distributed as D
        self.args, self.rank, self.world, self.dev = args, rank, world, dev
        self.D, self.ffi, self.torch = D, _ffi, torch
        self.sc, self.desc = scene_for(args, name)
        self.name = name or args.workload
        self.cam = self.sc.camera.to_c()
        self.W, self.H, self.spp = self.cam.screen_width, self.cam.screen_height, self.cam.aa_sample_count
        t0 = time.perf_counter()
        self.g = self.sc.commit(local)
        self.build_s = time.perf_counter() - t0
        self.shard = shard or args.shard
        self.accum = D.new_accum(self.W, self.H, dev)
        self.total_spp = self.spp * world if (world > 1 and self.shard == "weak") else self.spp

    def opts(self, flags=0, engine=None, sample_end=0):
        a, D, F = self.args, self.D, self.ffi
        kw = dict(wavefront=a.wavefront, flags=flags, engine=F.ENGINES[engine or a.engine],
                  ray_sort={"auto": 0, "off": 1, "on": 2}[a.ray_sort], blocks_per_sm=a.blocks_per_sm, tile=a.tile,
                  work_order={"auto": 0, "pixel": 1, "sample": 2, "grouped": 3}[a.work_order],
                  sample_end=sample_end)
        if a.emulate_shards:
            return D.shard_opts(a.emulate_rank, a.emulate_shards, SEED, self.shard, **kw)
        if self.world == 1 or self.shard == "weak":
            return D.shard_opts(0, 1, SEED + (self.rank if self.shard == "weak" else 0), "all", **kw)
        return D.shard_opts(self.rank, self.world, SEED, self.shard, **kw)

    def step(self, opts, collect=None, resolve_spp=None):
        D = self.D
        self.accum.zero_()
        st = D.render_shard(self.g, self.cam, opts, self.accum)
        D.reduce_accum(self.accum, dst=0)
        out = None
        if self.rank == 0:
            out = D.resolve(self.g, self.cam, self.accum, resolve_spp or self.total_spp)
        if collect is not None:
            for k, v in st.as_dict().items():
                collect[k] = collect.get(k, 0) + v
        return out

    def close(self):
        self.accum = None
        self.g.close()
        self.torch.cuda.empty_cache()


def timed_steps(job, steps, barrier, flush, clocks=None):
    """K steps bracketed by a barrier + synchronize on both sides, CUDA events on the render stream, max over ranks."""
    import torch
    import torch.distributed as dist
    stats = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if clocks is not None:
        clocks.start()
    e0.record()
    for _ in range(steps):
        flush.zero_()                 # evict the previous frame from L2 between timed iterations
        job.step(job.opts(), stats)
    e1.record()
    barrier()
    clock_info = clocks.stop() if clocks is not None else None
    ms = e0.elapsed_time(e1)
    keys = ("samples", "rays", "kernel_launches")
    if job.world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=job.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([stats[k] for k in keys], dtype=torch.float64, device=job.dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        totals = dict(zip(keys, (float(x) for x in tot.tolist())))
    else:
        totals = {k: float(stats[k]) for k in keys}
    return ms, totals, stats, clock_info


def config_entry(args, name, rank, world, local, dev, barrier, flush, shard=None, with_cpu=True):
    """One line of the `configs` block: the named BASELINE configuration at its full size on the final kernels."""
    job = Job(args, name, rank, world, local, dev, shard=shard)
    big = job.W * job.H * job.spp > (1 << 32)                       # C5: one frame is tens of seconds on one GPU
    warm_spp = 64 if big else 0                                     # warm-up on a window of the sample indices
    t0 = time.perf_counter()
    for _ in range(1 if big else 2):
        job.step(job.opts(sample_end=warm_spp), resolve_spp=warm_spp or None)
    job.torch.cuda.synchronize()
    steps = 1 if big else 3
    if not big:
        # a frame of C1 takes 7 ms: three of them end before the clocks have come back up from the idle stretch of the
        # previous configuration's CPU baseline.  Warm up for at least 0.3 s and time at least 0.5 s of frames.
        per = max((time.perf_counter() - t0) / 2, 1e-4)
        for _ in range(min(200, int(0.3 / per))):
            job.step(job.opts())
        steps = max(3, min(200, int(0.5 / per) + 1))
        if world > 1:      # every rank must run the same number of steps
            n = job.torch.tensor([steps], dtype=job.torch.int64, device=dev)
            import torch.distributed as dist
            dist.broadcast(n, src=0)
            steps = int(n.item())
    ms, totals, stats, _ = timed_steps(job, steps, barrier, flush)
    entry = None
    if rank == 0:
        cam = job.sc.camera
        entry = {
            "workload": f"{name}: {job.desc}", "width": job.W, "height": job.H, "spp": job.spp, "path_depth": cam.path_depth,
            "n_gpus": world, "shard": "all" if world == 1 else job.shard, "steps": steps, "ms_per_step": ms / steps,
            "value": totals["samples"] / (ms * 1e-3) / 1e6, "unit": UNIT,
            "rays_per_sec_M": totals["rays"] / (ms * 1e-3) / 1e6, "rays_per_sample": totals["rays"] / max(totals["samples"], 1),
            "engine": ENGINE_NAMES.get(stats["engine"] // steps, "?"),
            "gpu_launches": int(totals["kernel_launches"]),
        }
        if with_cpu and not args.no_cpu_baseline and world == 1:
            entry["cpu_baseline"] = cpu_baseline(job.sc, max(3.0, args.cpu_seconds / 4))
            entry["speedup_vs_cpu_port"] = entry["value"] / entry["cpu_baseline"]["value"]
    job.close()
    return entry


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cs397raytracingsp22_b200 import _ffi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lib = _ffi.load()
    ndev = lib.rt_device_count()
    if ndev <= 0:
        raise SystemExit("bench.py: no CUDA device is usable and there is no CPU fallback "
                         f"({lib.rt_last_error().decode()})")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    job = Job(args, None, rank, world, local, dev)
    cam, W, H = job.cam, job.W, job.H
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # counted pass (untimed, wavefront engine: the device counters live there): same keys => same rays => same counts
    counted = {}
    if args.no_counters:
        counted = {k: 0 for k in _ffi.rt_stats().as_dict()}
    else:
        job.step(job.opts(flags=_ffi.RT_OPT_COUNTERS, engine="wavefront"), counted)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        job.step(job.opts())
    torch.cuda.synchronize()

    ms, totals, stats_acc, clock_info = timed_steps(job, args.steps, barrier, flush, ClockSampler(local) if rank == 0 else None)
    job_samples, job_rays, job_launches = totals["samples"], totals["rays"], totals["kernel_launches"]
    value = job_samples / (ms * 1e-3) / 1e6
    rays_M = job_rays / (ms * 1e-3) / 1e6

    # ---- end to end through the public call with HOST buffers (scene upload + render + read-back in the timed region)
    e2e = None
    if not args.no_e2e:
        lin_h = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
        rgb_h = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
        scene_bytes = job.g.device_bytes()
        lin_np, rgb_np = lin_h.numpy(), rgb_h.numpy()

        def e2e_step():
            job.g.upload()                               # host -> device: the lowered scene (rt_scene_upload)
            if world == 1:
                # the public host-buffer call, exactly what Scene::render_to_image's replacement makes: rt_render
                # renders, resolves and copies the linear and RGB8 images into the caller's (pinned) host buffers
                job.g.render(cam, job.opts(), out_linear=lin_np, out_rgb8=rgb_np)
                return
            out = job.step(job.opts())                   # N>1: rt_render_accum per rank + NCCL reduce + rt_resolve
            if rank == 0:
                lin_h.copy_(out[0], non_blocking=True)   # device -> host: linear radiance + RGB8 image
                rgb_h.copy_(out[1], non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 5))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": (job_samples / args.steps) * n_e2e / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(scene_bytes + 512) * world, "d2h_bytes_per_step": int(W * H * 3 * 5),
               "steps": n_e2e,
               "call": ("rt_scene_upload + rt_render (host buffers in, host images out)" if world == 1 else
                        "rt_scene_upload + rt_render_accum per rank + NCCL reduce + rt_resolve + copy to pinned host buffers")}

    # ---- the other shard mode beside the main line (N>1), and the other BASELINE configurations
    beside = None
    if world > 1 and args.shard in ("tiles", "samples") and not args.no_configs:
        other = "samples" if args.shard == "tiles" else "tiles"
        j2 = Job(args, None, rank, world, local, dev, shard=other)
        for _ in range(2):
            j2.step(j2.opts())
        ms2, tot2, _, _ = timed_steps(j2, max(1, min(args.steps, 5)), barrier, flush)
        beside = {"shard": other, "value": tot2["samples"] / (ms2 * 1e-3) / 1e6, "unit": UNIT,
                  "ms_per_step": ms2 / max(1, min(args.steps, 5))}
        j2.close()
    configs = None
    default_job = not (args.spp or args.width or args.height or args.depth or args.emulate_shards)
    if not args.no_configs and default_job:
        configs = {}
        if world == 1:
            for name in ("c1", "c2", "c3", "c5"):
                if name != args.workload:
                    configs[name] = config_entry(args, name, rank, world, local, dev, barrier, flush)
        else:
            for mode in ("tiles", "samples"):
                configs[f"c5_{mode}"] = config_entry(args, "c5", rank, world, local, dev, barrier, flush, shard=mode)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel, from CUDA events recorded around every launch in the timed region.
    # Wavefront engine: k_trace, one launch per iteration.  Megakernel: k_path, one launch per frame.
    n_ext = max(stats_acc["extend_launches"], 1)
    ext_ms = stats_acc["ms_extend"] / n_ext
    megakernel = ENGINE_NAMES.get(stats_acc["engine"] // args.steps) == "megakernel"
    kernel = "k_path" if megakernel else "k_trace"
    alg_bytes = trace_algorithmic_bytes(counted) / (1 if megakernel else max(counted["extend_launches"], 1))
    if megakernel:
        alg_bytes -= 68.0 * counted["rays"]      # no ray queue / hit records in the megakernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "kernel_ncu.json"))).get(f"{job.name}:{kernel}", {})
    except Exception:
        pass
    rays_per_launch = stats_acc["rays"] / n_ext      # this rank's rays and launches
    roofline = {
        "kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": prof.get("dram_bytes_per_launch"),
        "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
        "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ext_ms, "launches": int(n_ext),
        "share_of_step": stats_acc["ms_extend"] / max(stats_acc["ms_total"], 1e-9),
        "shade_share_of_step": stats_acc["ms_shade"] / max(stats_acc["ms_total"], 1e-9),
        "bytes_per_ray": trace_algorithmic_bytes(counted) / max(counted["rays"], 1),
        "nodes_per_ray": counted["nodes_visited"] / max(counted["rays"], 1),
        "tlas_nodes_per_ray": counted["tlas_nodes_visited"] / max(counted["rays"], 1),
        "tris_per_ray": counted["tris_tested"] / max(counted["rays"], 1),
        "traversal_simt_efficiency": counted["nodes_visited"] / max(counted["warp_node_slots"], 1),
        "note": "SURVEY.md §8d accounting.  The lowered BVH fits in L1/L2, so these bytes are served on chip and the "
                "fraction can exceed 1: HBM is NOT the ceiling that binds this kernel - see roofline_issue",
    }
    # the ceiling that binds: issue slots.  thread-instructions per ray come from one ncu capture of steady-state
    # launches of this kernel on this workload (profiles/kernel_ncu.json); the rate is formed with the live timings.
    roofline_issue = None
    if prof.get("thread_inst_per_ray"):
        sm_mhz = (clock_info or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        peak_tips = B200_SMS * LANES_PER_SM_CLK * sm_mhz * 1e6 / 1e12
        ach_tips = prof["thread_inst_per_ray"] * rays_per_launch / (ext_ms * 1e-3) / 1e12
        roofline_issue = {
            "kernel": kernel, "bound": "issue", "achieved": ach_tips, "peak": peak_tips, "unit": "T thread-instructions/s",
            "frac": ach_tips / peak_tips,
            "peak_source": f"{B200_SMS} SMs x 4 schedulers x 32 lanes x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)",
            "thread_inst_per_ray": prof["thread_inst_per_ray"], "rays_per_launch": rays_per_launch,
            "ncu": {k: v for k, v in prof.items() if k not in ("thread_inst_per_ray", "dram_bytes_per_launch")},
        }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(job.sc, args.cpu_seconds)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": scaling_of(args),
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, job.sc, job.desc),
        "engine": "megakernel" if megakernel else "wavefront",
        "wavefront": int(args.wavefront or (1 << 25)), "scene_build_s": job.build_s, "scene_bytes": int(job.g.device_bytes()),
        "rays_per_sec_M": rays_M, "rays_per_sample": job_rays / max(job_samples, 1),
        "gpu_launches": int(job_launches),
        "clocks": clock_info,
        "e2e": e2e, "roofline": roofline, "roofline_issue": roofline_issue, "cpu_baseline": cpu,
    }
    if beside:
        line["beside"] = beside
    if configs:
        line["configs"] = configs
    if cpu:
        line["speedup_vs_cpu_port"] = {"device_resident": value / cpu["value"], "e2e": (e2e["value"] / cpu["value"]) if e2e else None}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
