#!/bin/bash
# round 2, GPU call 22 (1 GPU): ray-sort key space with 32 Mi rays in flight: 2^17 / 2^18 (base) / 2^19 / 2^20 bins, 32 or 64 cells per axis
mkdir -p gpurun_out
{
echo "== ray-sort bins 2^18 x 32 cells (base) / 2^19 x 32 / 2^19 x 64 / 2^20 x 64 / 2^17 x 32; 256 spp"
timeout 1500 python tools/ab.py build/rt_base.so build/rt_bins19.so build/rt_bins19c64.so build/rt_bins20c64.so build/rt_bins17.so -- "--workload c4 --engine wavefront" "--workload c5 --engine wavefront"
} > gpurun_out/r2_ab22.log 2>&1
cut -c1-215 gpurun_out/r2_ab22.log
