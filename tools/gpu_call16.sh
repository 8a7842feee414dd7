#!/bin/bash
# round 2, GPU call 16 (1 GPU): block size of the ray kernels (64 / 128 / 256 threads at the same warps per SM), four
# stack entries in shared memory on top of the packed pairs, a 32 Mi wavefront
mkdir -p gpurun_out
{
echo "== RT_BLOCK 128 (base) / 64 / 256, RT_SMEM_STACK 4; 256 spp"
timeout 1500 python tools/ab.py build/rt_base.so build/rt_b64.so build/rt_b256.so build/rt_s4.so -- "--workload c4 --engine wavefront" "--workload c3" "--workload c1"
echo "== wavefront of 32 Mi paths (default 16 Mi), C4 at full size"
timeout 600 python tools/ab.py build/rt_base.so -- "--full --no-counters" "--full --no-counters --wavefront 33554432"
} > gpurun_out/r2_ab16.log 2>&1
cut -c1-215 gpurun_out/r2_ab16.log
RT_B200_LIB=build/rt_b64.so timeout 600 python -m pytest tests/test_gpu_engines.py tests/test_gpu_parity.py -x -q -k "megakernel_equals or primary or low_spp or sharding" 2>&1 | tail -2
