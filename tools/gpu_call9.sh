#!/bin/bash
# round 2, GPU call 9 (1 GPU): margins of the image-parity tolerances; L1 / L2 experiments on k_trace: shared-memory
# stack depth (the carve-out it implies), L2 persistence window on nodes + triangles, L1 eviction hints
mkdir -p gpurun_out
timeout 600 python tools/parity_margins.py > gpurun_out/r2_parity_margins.log 2>&1; echo "margins rc=$?"
cat gpurun_out/r2_parity_margins.log
{
echo "== C4, 256 spp: stack depth in shared memory 16 (base) / 7 / 4 / 0, L2 persistence, L1 hints"
timeout 1500 python tools/ab.py build/rt_base.so build/rt_stack7.so build/rt_stack4.so build/rt_stack0.so build/rt_l2p.so build/rt_l2p7.so build/rt_nodelast.so build/rt_trina.so build/rt_hints7.so -- "--workload c4 --engine wavefront"
echo "== C5, 256 spp"
timeout 900 python tools/ab.py build/rt_base.so build/rt_stack7.so build/rt_l2p.so build/rt_hints7.so -- "--workload c5 --engine wavefront"
} > gpurun_out/r2_ab9.log 2>&1
cut -c1-215 gpurun_out/r2_ab9.log
