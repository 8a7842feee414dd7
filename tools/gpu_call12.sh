#!/bin/bash
# round 2, GPU call 12 (1 GPU): packed child pairs (two centres + links + six bf16 half extents = three 128-bit fetches per
# visit instead of four) against the base and the unpacked centre / half-extent form; GPU suite on the packed build
mkdir -p gpurun_out
{
echo "== base (min, max; 4 LDG.128) / ch (centre, half extent; 4 LDG.128) / ch2 (packed; 3 LDG.128), 256 spp"
timeout 1500 python tools/ab.py build/rt_base.so build/rt_ch.so build/rt_ch2.so -- "--workload c4 --engine wavefront" "--workload c5 --engine wavefront" "--workload c2"
} > gpurun_out/r2_ab12.log 2>&1
cut -c1-215 gpurun_out/r2_ab12.log
RT_B200_LIB=build/rt_ch2.so timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu12_ch2.log 2>&1; echo "gpu suite on ch2 rc=$?"
tail -6 gpurun_out/r2_pytest_gpu12_ch2.log
