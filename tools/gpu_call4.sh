#!/bin/bash
# round 2, GPU call 4 (2 GPUs): whole GPU suite on the current defaults, then bench at N=1 (short) and N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "gpu suite rc=$?"
tail -4 gpurun_out/r2_pytest_gpu4.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-configs > gpurun_out/bench_r2b_1gpu.json 2> gpurun_out/bench_r2b_1gpu.err; echo "bench N=1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r2b_1gpu.json')); print({k: d[k] for k in ('value','ms_per_step','engine')}, d['e2e']['value'], d['roofline_issue'] and d['roofline_issue']['frac'], d['roofline']['traffic'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2b_2gpu.json 2> gpurun_out/bench_r2b_2gpu.err; echo "bench N=2 rc=$?"
tail -3 gpurun_out/bench_r2b_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r2b_2gpu.json')); print({k: d[k] for k in ('value','ms_per_step','engine','scaling','n_gpus')}, 'e2e', d['e2e']['value'], 'beside', d.get('beside')); print({k:(v['value'],v['ms_per_step']) for k,v in d.get('configs',{}).items()})"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2b_2gpu_ref.json 2>/dev/null; echo "ref N=2 rc=$?"; cut -c1-300 gpurun_out/bench_r2b_2gpu_ref.json
