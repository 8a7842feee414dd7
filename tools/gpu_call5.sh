#!/bin/bash
# round 2, GPU call 5 (8 GPUs): the driver's SCALE command at N=8 - one C4 frame strong-scaled by tiles (samples beside it),
# C5 at full size by tiles and by samples
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2c_8gpu.json 2> gpurun_out/bench_r2c_8gpu.err; echo "bench N=8 rc=$?"
tail -3 gpurun_out/bench_r2c_8gpu.err
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2c_8gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print({k: d[k] for k in ('value','ms_per_step','engine','scaling','n_gpus')}, 'e2e', d['e2e']['value'], 'beside', d.get('beside'))
        print({k:(v['value'],v['ms_per_step']) for k,v in d.get('configs',{}).items()})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 --no-configs > gpurun_out/bench_r2c_4gpu.json 2> gpurun_out/bench_r2c_4gpu.err; echo "bench N=4 rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2c_4gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print({k: d[k] for k in ('value','ms_per_step','engine','scaling','n_gpus')}, 'e2e', d['e2e']['value'], 'beside', d.get('beside'))
PY
