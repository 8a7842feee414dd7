#!/bin/bash
# round 2, GPU call 24 (1 GPU): bench.py with default flags, as the driver runs it; wall-clock of the whole command
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r2g_default.json 2> gpurun_out/bench_r2g_default.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"
python -c "
import json; d=json.load(open('gpurun_out/bench_r2g_default.json')); print({k: d[k] for k in ('value','ms_per_step','steps','warmup','engine','gpu_launches')}, d['e2e']['value'], d['roofline_issue']['frac'], d['roofline']['traffic'], d['roofline']['frac']); print({k:(round(v['value'],1),v['engine']) for k,v in d['configs'].items()}); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'])"
