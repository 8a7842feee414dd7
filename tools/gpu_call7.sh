#!/bin/bash
# round 2, GPU call 7: GPU suite on the new defaults (ranked sort, diagonal tiles, finer polling, parallel BVH build),
# the eight tile shards of a C4 frame one after the other on one GPU (balance), bench N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu7.log 2>&1; echo "gpu suite rc=$?"
tail -4 gpurun_out/r2_pytest_gpu7.log
{
echo "== the eight shards of one C4 frame, one GPU: diagonal 16x16 tiles, and sample ranges for comparison"
for r in 0 1 2 3 4 5 6 7; do
  timeout 300 python tools/ab.py cs397raytracingsp22_b200/librt_b200.so -- "--full --no-counters --emulate-shards 8 --emulate-rank $r --shard tiles"
done
timeout 300 python tools/ab.py cs397raytracingsp22_b200/librt_b200.so -- "--full --no-counters --emulate-shards 8 --emulate-rank 0 --shard samples" "--full --no-counters --emulate-shards 8 --emulate-rank 5 --shard samples" "--full --no-counters --emulate-shards 8 --emulate-rank 3 --shard tiles --tile 32"
} > gpurun_out/r2_ab7.log 2>&1
cut -c1-170 gpurun_out/r2_ab7.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_r2d_1gpu.json 2> gpurun_out/bench_r2d_1gpu.err; echo "bench N=1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r2d_1gpu.json')); print({k: d[k] for k in ('value','ms_per_step','engine','gpu_launches')}, d['e2e']['value'], d['roofline_issue']['frac'], d['scene_build_s']); print({k:(round(v['value'],1),v['engine']) for k,v in d['configs'].items()})"
