#!/bin/bash
# round 2, GPU call 19 (8 GPUs): the driver's SCALE command at N=8 on the final kernels - one C4 frame strong-scaled by
# tiles (samples beside it), C5 at full size by tiles and by samples; then N=4 on four of the GPUs
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/bench_r2f_8gpu.err | grep '^{' > gpurun_out/bench_r2f_8gpu.json; echo "bench N=8 rc=${PIPESTATUS[0]}"
tail -2 gpurun_out/bench_r2f_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 4 --steps 20 --warmup 5 2> gpurun_out/bench_r2f_4gpu.err | grep '^{' > gpurun_out/bench_r2f_4gpu.json; echo "bench N=4 rc=${PIPESTATUS[0]}"
python - <<'PY'
import json
for n in (8, 4):
    d = json.load(open(f'gpurun_out/bench_r2f_{n}gpu.json'))
    print({k: d[k] for k in ('value', 'ms_per_step', 'engine', 'scaling', 'n_gpus')}, 'e2e', d['e2e']['value'], 'beside', d.get('beside'))
    print({k: (v['value'], v['ms_per_step']) for k, v in d.get('configs', {}).items()})
PY
