#!/usr/bin/env python
"""Fill profiles/kernel_ncu.json (what bench.py's roofline_issue and roofline.traffic read) from the raw page of a
steady-state ncu capture made by profiles/summarize.py.

usage: python profiles/make_kernel_ncu.py profiles/NAME_raw.csv WORKLOAD RAYS_PER_LAUNCH "source description"
       e.g. ... profiles/r2e_k_trace_k_shade_steady_raw.csv c4 16777216 "profiles/r2e_...md: ncu --set full ..."
Every kernel in the capture gets an entry "WORKLOAD:kernel" (k_trace, k_shade, k_path ...); the per-launch lists keep
the order of the capture.  Steady-state launches of the wavefront engine hold exactly `capacity` rays (the wavefront
is refilled to capacity every iteration), which is what RAYS_PER_LAUNCH states.
"""
import csv
import json
import os
import re
import sys

STALLS = ["long_scoreboard", "wait", "barrier", "not_selected", "math_pipe_throttle", "branch_resolving", "short_scoreboard",
          "no_instruction", "dispatch_stall", "lg_throttle", "mio_throttle"]


def main():
    raw, workload, rays, source = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,                      # durations -> ms
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}  # sizes -> bytes

    def f(r, name):
        return float(r[col[name]].replace(",", "")) * scale.get(units[col[name]], 1.0)

    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_ncu.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    kernels = {}
    for r in data:
        m = re.search(r"(k_[a-z_]+)", r[col["Kernel Name"]])
        if m:
            kernels.setdefault(m.group(1), []).append(r)
    for k, rs in kernels.items():
        warp_inst = [f(r, "smsp__inst_executed.sum") for r in rs]
        lanes = [f(r, "smsp__thread_inst_executed_per_inst_executed.ratio") for r in rs]
        stall = {s: sum(f(r, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio") for r in rs) / len(rs) for s in STALLS}
        out[f"{workload}:{k}"] = {
            "source": source,
            "launch_duration_ms": [f(r, "gpu__time_duration.sum") for r in rs],
            "thread_inst_per_ray": sum(w * l for w, l in zip(warp_inst, lanes)) / (rays * len(rs)),
            "warp_instructions_per_launch": warp_inst,
            "active_threads_per_instruction": lanes,
            "issue_active_pct": [f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") for r in rs],
            "alu_pipe_pct": [f(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") for r in rs],
            "fma_pipe_pct": [f(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") for r in rs],
            "l1_data_pipe_pct": [f(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") for r in rs],
            "l1_sector_hit_pct": [f(r, "l1tex__t_sector_hit_rate.pct") for r in rs],
            "l2_sector_hit_pct": [f(r, "lts__t_sector_hit_rate.pct") for r in rs],
            "l1tex_throughput_pct": [f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed") for r in rs],
            "dram_throughput_pct": [f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") for r in rs],
            "dram_bytes_per_launch": sum(f(r, "dram__bytes_read.sum") + f(r, "dram__bytes_write.sum") for r in rs) / len(rs),
            "top_stalls": [s for s, _ in sorted(stall.items(), key=lambda kv: -kv[1])[:3]],
        }
    json.dump(out, open(out_path, "w"), indent=1)
    for k in kernels:
        e = out[f"{workload}:{k}"]
        print(k, "thread-inst/ray %.1f" % e["thread_inst_per_ray"], "dram bytes/launch %.3g" % e["dram_bytes_per_launch"], e["top_stalls"])


if __name__ == "__main__":
    main()
