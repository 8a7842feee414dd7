#!/bin/bash
# round 2, GPU call 25 (1 GPU): small knobs once more on the final defaults: ray indices per claim (32 / 64 / 128), stack
# entries in shared memory (4 / 0 / 8), triangles per leaf (4 / 2 / 8)
mkdir -p gpurun_out
{
echo "== base / claim 64 / claim 128 / stack 0 / stack 8 / leaf 2 / leaf 8; 256 spp"
timeout 1700 python tools/ab.py build/rt_base.so build/rt_chunk64.so build/rt_chunk128.so build/rt_stack0.so build/rt_stack8.so build/rt_leaf2.so build/rt_leaf8.so -- "--workload c4 --engine wavefront" "--workload c5 --engine wavefront"
} > gpurun_out/r2_ab25.log 2>&1
cut -c1-215 gpurun_out/r2_ab25.log
