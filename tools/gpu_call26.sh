#!/bin/bash
# round 2, GPU call 26 (1 GPU): 64 ray indices per claim, alone and with the whole stack (or all but 2 entries) in local memory
mkdir -p gpurun_out
{
echo "== base / claim 64 / claim 64 + stack 0 / claim 64 + stack 2; 256 spp, then C4 at full size"
timeout 1700 python tools/ab.py build/rt_base.so build/rt_c64.so build/rt_c64s0.so build/rt_c64s2.so -- "--workload c4 --engine wavefront" "--workload c5 --engine wavefront" "--full --no-counters"
} > gpurun_out/r2_ab26.log 2>&1
cut -c1-215 gpurun_out/r2_ab26.log
