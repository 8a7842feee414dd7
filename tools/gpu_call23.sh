#!/bin/bash
# round 2, GPU call 23 (1 GPU): the committed state as the driver will run it - GPU suite, smoke(), bench with default flags
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu23.log 2>&1; echo "gpu suite rc=$?"
tail -2 gpurun_out/r2_pytest_gpu23.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/bench_r2g_default.json 2> gpurun_out/bench_r2g_default.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/bench_r2g_default.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2g_default.json')); print({k: d[k] for k in ('value','ms_per_step','steps','warmup','engine','gpu_launches')}, d['e2e']['value'], d['roofline_issue']['frac'], d['roofline']['traffic'], d['roofline']['frac']); print({k:(round(v['value'],1),v['engine']) for k,v in d['configs'].items()}); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
