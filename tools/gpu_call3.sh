#!/bin/bash
# round 2, GPU call 3: megakernel occupancy variants, early refill on the closed scenes, k_shade warp scan, C5 work order,
# steady-state ncu capture (no counted pass), GPU suite on the new defaults
mkdir -p gpurun_out
{
echo "== megakernel: 8 / 7 / 6 resident blocks (64 / 72 / 76 registers)"
timeout 900 python tools/ab.py build/rt_base.so build/rt_mk7.so build/rt_mk6.so -- "--workload c1 --engine megakernel" "--workload c2 --engine megakernel" "--workload c3 --engine megakernel"
echo "== wavefront: k_trace early refill (idle lanes that trigger a refill; 32 = whole warp) on the closed scenes"
timeout 900 python tools/ab.py build/rt_base.so build/rt_refill16.so build/rt_refill8.so build/rt_refill4.so -- "--workload c1 --engine wavefront" "--workload c2 --engine wavefront" "--workload c3 --engine wavefront"
echo "== wavefront: k_shade warp scan"
timeout 600 python tools/ab.py build/rt_base.so build/rt_noscan.so -- "--workload c1 --engine wavefront" "--workload c4 --engine wavefront" "--workload c3 --engine wavefront"
echo "== C5 full size: work order"
timeout 900 python tools/ab.py build/rt_base.so -- "--workload c5 --full --steps 1 --warmup 1 --no-counters --work-order pixel" "--workload c5 --full --steps 1 --warmup 1 --no-counters --work-order grouped"
echo "== C4 full size: work order"
timeout 600 python tools/ab.py build/rt_base.so -- "--workload c4 --full --work-order pixel" "--workload c4 --full --work-order grouped"
} > gpurun_out/r2_ab3.log 2>&1
cut -c1-200 gpurun_out/r2_ab3.log
echo "== steady-state ncu capture (k_trace / k_shade launches from the middle of a frame)"
BENCH="python bench.py --steps 1 --warmup 3 --spp 256 --no-cpu-baseline --no-e2e --no-configs --no-counters"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -s 60 -c 4 -o gpurun_out/prof_r2b -f $BENCH > gpurun_out/ncu_r2b.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 420 --csv --log-file gpurun_out/launches_r2b.csv $BENCH > gpurun_out/ncu_r2b_list.log 2>&1; echo "ncu list rc=$?"
echo "== megakernel ncu capture (C1)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_path" -s 3 -c 1 -o gpurun_out/prof_r2b_kpath -f python bench.py --workload c1 --engine megakernel --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --no-counters > gpurun_out/ncu_r2b_kpath.log 2>&1; echo "ncu kpath rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "gpu suite rc=$?"
tail -5 gpurun_out/r2_pytest_gpu3.log
