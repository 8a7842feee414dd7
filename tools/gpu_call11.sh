#!/bin/bash
# round 2, GPU call 11 (1 GPU): is the node visit of k_trace bound by L1 data-pipe wavefronts?  1 / 2 / 4 extra 128-bit
# fetches per visit (no extra arithmetic) against the base
mkdir -p gpurun_out
{
echo "== k_trace with 0 / 1 / 2 / 4 extra LDG.128 per node visit (4 wavefronts each on top of the visit's 16), C4 256 spp"
timeout 900 python tools/ab.py build/rt_base.so build/rt_x1.so build/rt_x2.so build/rt_x4.so -- "--workload c4 --engine wavefront"
} > gpurun_out/r2_ab11.log 2>&1
cut -c1-215 gpurun_out/r2_ab11.log
