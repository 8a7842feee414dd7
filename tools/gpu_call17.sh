#!/bin/bash
# round 2, GPU call 17 (1 GPU): wavefront width 16 / 32 / 64 Mi on whole frames and on one eighth of a frame, with and
# without 256-thread blocks + 4 stack entries in shared memory
mkdir -p gpurun_out
{
echo "== C4 full size: wavefront 16 / 32 / 64 Mi; base and b256s4"
timeout 1500 python tools/ab.py build/rt_base.so build/rt_b256s4.so -- "--full --no-counters" "--full --no-counters --wavefront 33554432" "--full --no-counters --wavefront 67108864"
echo "== one eighth of the C4 frame (tile shard 3 of 8): wavefront 16 / 32 Mi"
timeout 900 python tools/ab.py build/rt_base.so build/rt_b256s4.so -- "--full --no-counters --emulate-shards 8 --emulate-rank 3 --shard tiles" "--full --no-counters --emulate-shards 8 --emulate-rank 3 --shard tiles --wavefront 33554432" "--full --no-counters --emulate-shards 8 --emulate-rank 3 --shard samples --wavefront 33554432"
echo "== C5 full size (4096 spp), one frame: wavefront 16 / 32 Mi"
timeout 900 python tools/ab.py build/rt_base.so -- "--workload c5 --full --steps 1 --warmup 1 --no-counters" "--workload c5 --full --steps 1 --warmup 1 --no-counters --wavefront 33554432"
} > gpurun_out/r2_ab17.log 2>&1
cut -c1-215 gpurun_out/r2_ab17.log
