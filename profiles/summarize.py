#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full --import-source on) into the markdown summary committed under profiles/.
usage: python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/NAME   -> NAME.md, NAME_raw.csv"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sectors.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_op_read_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
]
STALLS = ["long_scoreboard", "wait", "barrier", "not_selected", "math_pipe_throttle", "branch_resolving", "short_scoreboard",
          "no_instruction", "dispatch_stall", "lg_throttle", "mio_throttle", "tex_throttle", "membar", "sleeping", "selected"]


def ncu(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = ncu(["--page", "raw", "--csv"])
    open(out + "_raw.csv", "w").write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    md = [f"# ncu summary of `{rep.split('/')[-1]}`\n",
          "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (one GPU); per-launch times are "
          "cold-cache and serialised, so compare shares, not absolutes.  Raw page: `" + out.split('/')[-1] + "_raw.csv`.\n"]
    md.append("| metric | unit | " + " | ".join(r[name_i].split("(")[0].replace("void ", "").replace("rt::", "") for r in data) + " |")
    md.append("|---|---|" + "---|" * len(data))
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            md.append(f"| `{m}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    md.append("\nWarp stall reasons (warps stalled per issue-active cycle):\n")
    md.append("| stall | " + " | ".join(str(k) for k in range(len(data))) + " |")
    md.append("|---|" + "---|" * len(data))
    for s in STALLS:
        m = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if m in hdr:
            i = hdr.index(m)
            md.append(f"| {s} | " + " | ".join(f"{float(r[i]):.2f}" for r in data) + " |")
    # hottest source lines of the first two kernels
    src = ncu(["--page", "source", "--csv", "--print-source", "cuda,sass"])
    rows = list(csv.reader(io.StringIO(src)))
    fstarts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
    seen = set()
    for si in fstarts:
        fn = rows[si][1].split("(")[0]
        if fn in seen:
            continue
        seen.add(fn)
        h = rows[si + 1]
        end = next((i for i in range(si + 2, len(rows)) if rows[i] and rows[i][0] in ("Line No", "File Path", "Function Name")), len(rows))
        iI, iT, iS = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
        agg = []
        for r in rows[si + 2:end]:
            if len(r) <= iT or r[2] != "-":
                continue
            try:
                agg.append((int(r[0]), r[1].strip()[:96], int(r[iI]), int(r[iT]), int(r[iS])))
            except ValueError:
                pass
        tI, tT, tS = sum(a[2] for a in agg), sum(a[3] for a in agg), sum(a[4] for a in agg)
        if tI < 1_000_000:
            continue
        md.append(f"\n## hottest source lines — `{fn}` (first captured launch)\n")
        md.append(f"{tI} warp instructions, {tT / max(tI, 1):.1f} active threads per instruction on average, {tS} stall samples.\n")
        md.append("| line | % of instructions | active threads | % of stall samples | source |")
        md.append("|---|---|---|---|---|")
        for a in sorted(agg, key=lambda a: -a[2])[:18]:
            md.append(f"| {a[0]} | {100 * a[2] / tI:.1f} | {a[3] / max(a[2], 1):.1f} | {100 * a[4] / max(tS, 1):.1f} | `{a[1].replace('|', '¦')}` |")
    open(out + ".md", "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
