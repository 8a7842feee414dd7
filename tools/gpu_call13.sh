#!/bin/bash
# round 2, GPU call 13 (1 GPU): the final kernels (packed child pairs) - GPU suite, smoke under compute-sanitizer,
# bench N=1 with the configs block, steady-state ncu capture and launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu13.log 2>&1; echo "gpu suite rc=$?"
tail -3 gpurun_out/r2_pytest_gpu13.log
# (a compute-sanitizer memcheck of smoke() stood here: the tool is closed on this GPU pool, gpurun refuses it)
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_r2e_1gpu.json 2> gpurun_out/bench_r2e_1gpu.err; echo "bench N=1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r2e_1gpu.json')); print({k: d[k] for k in ('value','ms_per_step','engine','gpu_launches')}, d['e2e']['value'], d['roofline_issue']['frac'], d['scene_build_s']); print({k:(round(v['value'],1),v['engine']) for k,v in d['configs'].items()}); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
BENCH="python bench.py --steps 1 --warmup 3 --spp 256 --no-cpu-baseline --no-e2e --no-configs --no-counters"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -s 60 -c 4 -o gpurun_out/prof_r2e -f $BENCH > gpurun_out/ncu_r2e.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 420 --csv --log-file gpurun_out/launches_r2e.csv $BENCH > gpurun_out/ncu_r2e_list.log 2>&1; echo "ncu list rc=$?"
