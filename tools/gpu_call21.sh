#!/bin/bash
# round 2, GPU call 21 (1 GPU): 512-thread blocks (k_shade at 2 or 3 resident blocks), k_sort block size
mkdir -p gpurun_out
{
echo "== base (256-thread blocks) / 512-thread blocks with k_shade at 2 (64 regs) or 3 (40 regs) blocks per SM / k_sort at 1024 or 256 threads; 256 spp"
timeout 1500 python tools/ab.py build/rt_base.so build/rt_b512a.so build/rt_b512b.so build/rt_sort1024.so build/rt_sort256.so -- "--workload c4 --engine wavefront" "--workload c3"
} > gpurun_out/r2_ab21.log 2>&1
cut -c1-215 gpurun_out/r2_ab21.log
