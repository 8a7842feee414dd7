#!/bin/bash
# round 2, GPU call 27 (1 GPU): final defaults - GPU suite, smoke(), bench N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu27.log 2>&1; echo "gpu suite rc=$?"
tail -2 gpurun_out/r2_pytest_gpu27.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_r2h_1gpu.json 2> gpurun_out/bench_r2h_1gpu.err; echo "bench N=1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r2h_1gpu.json')); print({k: d[k] for k in ('value','ms_per_step','engine','gpu_launches')}, d['e2e']['value'], d['roofline_issue']['frac'], d['roofline']['frac'], d['rays_per_sec_M']); print({k:(round(v['value'],1),v['engine'],v['steps']) for k,v in d['configs'].items()}); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['speedup_vs_cpu_port'])"
