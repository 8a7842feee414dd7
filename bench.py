#!/usr/bin/env python
"""bench.py — headline benchmark: Msamples/s (and Mrays/s) of the path-tracing hot path on B200.

  python bench.py --gpus 1 --steps K --warmup W            one GPU
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   one rank per GPU
  python bench.py --impl reference ...                     the reference-equivalent CPU renderer (oracle port)

A "step" is one full render of the workload (default: BASELINE config C4, the textured drone scene of the
reference's run(), 1920x1080, 1024 spp, depth 10).  With N ranks the scaling is WEAK: every rank renders the whole
frame at 1024 spp with its own Philox key (seed + rank), the exact int64 accumulators are summed with one NCCL
reduce and rank 0 resolves an image of N*1024 spp.  Prints ONE JSON line (rank 0).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--shard", default="weak", choices=["weak", "samples", "tiles"],
                    help="N>1: weak = every rank renders the full spp with its own key (default); "
                         "samples/tiles = strong scaling of one frame")
    ap.add_argument("--wavefront", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (invalidates the headline)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--emulate-shards", type=int, default=0, help="diagnostics: render shard 0 of this many (with --shard)")
    ap.add_argument("--depth", type=int, default=0, help="override path_depth (diagnostics; invalidates the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def scene_for(args):
    from cs397raytracingsp22_b200 import scenes
    kw = {}
    if args.spp:
        kw["spp"] = args.spp
    if args.width:
        kw["width"] = args.width
    if args.height:
        kw["height"] = args.height
    if args.depth:
        kw["depth"] = args.depth
    return scenes.make_scene(args.workload, **kw), scenes.DESCRIPTIONS[args.workload]


def workload_config(args, sc, desc, extra=None):
    cam = sc.camera
    cfg = {
        "workload": f"{args.workload}: {desc}",
        "width": cam.screen_width, "height": cam.screen_height, "spp": cam.aa_sample_count,
        "path_depth": cam.path_depth, "objects": len(sc.objects),
        "seed": SEED,
    }
    if extra:
        cfg.update(extra)
    return cfg


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
    the full-resolution frame for sample indices [first, first+n_spp).  Returns (samples, rays, seconds, cores)."""
    import oracle_ffi as O
    b = O.lower_to_oracle(sc)
    cam = sc.camera.to_c()
    cores = O.load().orc_num_threads()
    t0 = time.perf_counter()
    _, _, st = b.render(cam, seed=SEED, mode=O.MODE_REF_TREE, sample_begin=first, sample_end=first + n_spp,
                        nthreads=0, want_linear=True, want_rgb8=False)
    dt = time.perf_counter() - t0
    b.close()
    return int(st.samples), int(st.rays), dt, cores


def cpu_baseline(sc, target_seconds: float) -> dict:
    s, r, dt, cores = cpu_render_sample(sc, 1)
    n = max(1, min(int(target_seconds / max(dt, 1e-3)) - 1, sc.camera.aa_sample_count - 1, 16))
    if n >= 1 and dt < target_seconds * 0.6:
        s2, r2, dt2, _ = cpu_render_sample(sc, n, first=1)
        s, r, dt = s + s2, r + r2, dt + dt2
    cam = sc.camera
    return {"value": s / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
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
              f"indices per pixel ({cam.screen_width * cam.screen_height} paths), oracle port in reference-tree mode")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_t / max(args.steps, 1) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sc, desc),
        "rays_per_sec_M": tot_r / tot_t / 1e6,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------ GPU arm
def trace_algorithmic_bytes(st) -> float:
    """Algorithmic bytes of ALL k_trace launches of one frame: 32 B per BVH node fetched,
    48 B per triangle record, 96 B of transforms per instance entered, 32 B per analytic primitive record, plus the
    wavefront's own traffic in this kernel: 48 B ray read + 20 B hit record written per ray."""
    return (32.0 * st["nodes_visited"] + 48.0 * st["tris_tested"] + 96.0 * st["instances_entered"]
            + 32.0 * st["prims_tested"] + 68.0 * st["rays"])


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cs397raytracingsp22_b200 import _ffi, distributed as D

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

    sc, desc = scene_for(args)
    cam_py = sc.camera
    cam = cam_py.to_c()
    W, H, spp = cam.screen_width, cam.screen_height, cam.aa_sample_count
    t0 = time.perf_counter()
    g = sc.commit(local)
    build_s = time.perf_counter() - t0

    def opts_for(flags=0):
        if args.emulate_shards:
            return D.shard_opts(0, args.emulate_shards, SEED, args.shard, wavefront=args.wavefront, flags=flags)
        if world == 1 or args.shard == "weak":
            return D.shard_opts(0, 1, SEED + (rank if args.shard == "weak" else 0), "all", wavefront=args.wavefront,
                                flags=flags)
        return D.shard_opts(rank, world, SEED, args.shard, wavefront=args.wavefront, flags=flags)

    total_spp = spp * world if (world > 1 and args.shard == "weak") else spp
    accum = D.new_accum(W, H, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stats_acc = {}

    def step(opts, collect=None):
        accum.zero_()
        st = D.render_shard(g, cam, opts, accum)
        D.reduce_accum(accum, dst=0)
        out = None
        if rank == 0:
            out = D.resolve(g, cam, accum, total_spp)
        if collect is not None:
            for k, v in st.as_dict().items():
                collect[k] = collect.get(k, 0) + v
        return out

    # counted pass (untimed): same keys => same counts as the timed passes
    counted = {}
    step(opts_for(_ffi.RT_OPT_COUNTERS), counted)
    torch.cuda.synchronize()
    for _ in range(max(args.warmup - 1, 0)):
        step(opts_for())
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        clocks.start()
    e0.record()
    for _ in range(args.steps):
        flush.zero_()                 # evict the previous frame from L2 between timed iterations
        step(opts_for(), stats_acc)
    e1.record()
    barrier()
    clock_info = clocks.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([stats_acc["samples"], stats_acc["rays"], stats_acc["kernel_launches"]], dtype=torch.float64,
                           device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        job_samples, job_rays, job_launches = (float(x) for x in tot.tolist())
    else:
        job_samples, job_rays, job_launches = float(stats_acc["samples"]), float(stats_acc["rays"]), float(
            stats_acc["kernel_launches"])
    value = job_samples / (ms * 1e-3) / 1e6
    rays_M = job_rays / (ms * 1e-3) / 1e6

    # ---- end to end through the public call with HOST buffers (scene upload + render + read-back in the timed region)
    e2e = None
    if not args.no_e2e:
        lin_h = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
        rgb_h = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
        scene_bytes = g.device_bytes()

        lin_np, rgb_np = lin_h.numpy(), rgb_h.numpy()

        def e2e_step():
            g.upload()                                   # host -> device: the lowered scene (rt_scene_upload)
            if world == 1:
                # the public host-buffer call, exactly what Scene::render_to_image's replacement makes: rt_render
                # renders, resolves and copies the linear and RGB8 images into the caller's (pinned) host buffers
                g.render(cam, opts_for(), out_linear=lin_np, out_rgb8=rgb_np)
                return
            out = step(opts_for())                       # N>1: rt_render_accum per rank + NCCL reduce + rt_resolve
            if rank == 0:
                lin_h.copy_(out[0], non_blocking=True)   # device -> host: linear radiance + RGB8 image
                rgb_h.copy_(out[1], non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": (job_samples / args.steps) * n_e2e / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(scene_bytes + 512), "d2h_bytes_per_step": int(W * H * 3 * 5),
               "steps": n_e2e,
               "call": ("rt_scene_upload + rt_render (host buffers in, host images out)" if world == 1 else
                        "rt_scene_upload + rt_render_accum per rank + NCCL reduce + rt_resolve + copy to pinned host buffers")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_trace), from CUDA events recorded around every launch in the timed region
    n_ext = max(stats_acc["extend_launches"], 1)
    ext_ms = stats_acc["ms_extend"] / n_ext
    alg_bytes = trace_algorithmic_bytes(counted) / max(counted["extend_launches"], 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    traffic, ncu_counters = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "k_trace_dram_traffic.json")))
        traffic = prof.get(args.workload)
        ncu_counters = prof.get(args.workload + "_ncu")   # what actually bounds the kernel (one ncu --set full capture)
    except Exception:
        pass
    roofline = {
        "kernel": "k_trace", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic,
        "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
        "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ext_ms, "launches": int(n_ext),
        "share_of_step": stats_acc["ms_extend"] / max(stats_acc["ms_total"], 1e-9),
        "shade_share_of_step": stats_acc["ms_shade"] / max(stats_acc["ms_total"], 1e-9),
        "bytes_per_ray": trace_algorithmic_bytes(counted) / max(counted["rays"], 1),
        "nodes_per_ray": counted["nodes_visited"] / max(counted["rays"], 1),
        "tlas_nodes_per_ray": counted["tlas_nodes_visited"] / max(counted["rays"], 1),
        "tris_per_ray": counted["tris_tested"] / max(counted["rays"], 1),
        "traversal_simt_efficiency": counted["nodes_visited"] / max(counted["warp_node_slots"], 1),
        "ncu": ncu_counters,
        "note": "the lowered scene fits in L2, so DRAM traffic is far below algorithmic bytes; the kernel is "
                "latency/issue bound (see profiles/)",
    }

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(sc, args.cpu_seconds)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if (world == 1 or args.shard == "weak") else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sc, desc, {
            "parallelism": f"{world} GPU(s), " + ("every rank renders the full frame at the configured spp with its own "
                                                  "Philox key; one NCCL int64 reduce per frame" if world > 1 and args.shard == "weak"
                                                  else ("single GPU" if world == 1 else f"one frame sharded by {args.shard}")),
            "total_spp": total_spp, "l2": "flush (256 MiB memset between timed iterations)",
            "wavefront": int(args.wavefront or (1 << 24)), "scene_build_s": build_s,
            "scene_bytes": int(g.device_bytes()),
        }),
        "rays_per_sec_M": rays_M, "rays_per_sample": job_rays / max(job_samples, 1),
        "gpu_launches": int(job_launches),
        "clocks": clock_info,
        "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
    }
    if cpu:
        line["speedup_vs_cpu_port"] = {"device_resident": value / cpu["value"], "e2e": (e2e["value"] / cpu["value"]) if e2e else None}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
