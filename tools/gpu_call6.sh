#!/bin/bash
# round 2, GPU call 6: occupancy of k_trace / k_shade, ranked ray sort
mkdir -p gpurun_out
{
echo "== k_trace at 8 / 9 / 10 resident blocks (64 / 56 / 48 registers); k_shade at 10 / 12 / 8 blocks; ranked sort"
timeout 1500 python tools/ab.py build/rt_base.so build/rt_trace9.so build/rt_trace10.so build/rt_shade12.so build/rt_shade8.so build/rt_ranked.so -- "--workload c4 --engine wavefront" "--workload c5 --engine wavefront" "--workload c3 --engine wavefront"
echo "== one eighth of the C4 frame on one GPU: where do tile shards lose 7 ms against sample shards?"
timeout 900 python tools/ab.py build/rt_base.so -- "--full --emulate-shards 8 --shard samples" "--full --emulate-shards 8 --shard tiles" "--full --emulate-shards 8 --shard tiles --work-order pixel" "--full --emulate-shards 8 --shard tiles --tile 64" "--full --emulate-shards 8 --shard tiles --tile 8" "--full --emulate-shards 8 --shard tiles --wavefront 8388608" "--full --emulate-shards 8 --shard samples --wavefront 8388608"
echo "== parity of the ranked sort and the 10-block build"
RT_B200_LIB=build/rt_ranked.so timeout 600 python -m pytest tests/test_gpu_engines.py -x -q -k "megakernel_equals" 2>&1 | tail -2
RT_B200_LIB=build/rt_trace10.so timeout 600 python -m pytest tests/test_gpu_engines.py tests/test_gpu_parity.py -x -q -k "megakernel_equals or primary or secondary" 2>&1 | tail -2
} > gpurun_out/r2_ab6.log 2>&1
cut -c1-215 gpurun_out/r2_ab6.log
