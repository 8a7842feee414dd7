#!/bin/bash
# round 2, GPU call 28 (2 GPUs): bench N=2 on the final kernels and the reference (CPU) arm as the driver launches it at N=2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench_r2h_2gpu.err | grep '^{' > gpurun_out/bench_r2h_2gpu.json; echo "bench N=2 rc=${PIPESTATUS[0]}"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/bench_r2h_2gpu_ref.json; echo "ref N=2 rc=${PIPESTATUS[0]}"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r2h_2gpu.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'engine', 'scaling', 'n_gpus')}, 'e2e', d['e2e']['value'], 'beside', d.get('beside'))
print({k: (v['value'], v['ms_per_step']) for k, v in d.get('configs', {}).items()})
r = json.load(open('gpurun_out/bench_r2h_2gpu_ref.json'))
print('reference arm:', r['value'], r['unit'], r['cpu_baseline']['cores'], 'cores;', 'same config keys:', r['config'] == d['config'])
PY
