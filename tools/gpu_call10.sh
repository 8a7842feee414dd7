#!/bin/bash
# round 2, GPU call 10 (1 GPU): node boxes as (centre, half extent) - slab test with 18 FFMA + 8 FMNMX instead of
# 12 FFMA + 20 FMNMX per child pair; A/B on C4 / C5 / C2 / C3 and the whole GPU suite on the variant
mkdir -p gpurun_out
{
echo "== (min, max) nodes [base] against (centre, half extent) nodes [ch], 256 spp"
timeout 1500 python tools/ab.py build/rt_base.so build/rt_ch.so -- "--workload c4 --engine wavefront" "--workload c5 --engine wavefront" "--workload c2" "--workload c3" "--workload c1"
} > gpurun_out/r2_ab10.log 2>&1
cut -c1-215 gpurun_out/r2_ab10.log
RT_B200_LIB=build/rt_ch.so timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu10_ch.log 2>&1; echo "gpu suite on ch rc=$?"
tail -6 gpurun_out/r2_pytest_gpu10_ch.log
