#!/bin/bash
# round 2, GPU call 2: 4-wide BVH parity + A/B, megakernel tuning, full default bench line, steady-state ncu capture
mkdir -p gpurun_out
{
echo "== parity of the 4-wide BVH build"
RT_B200_LIB=build/rt_bvh4.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engines.py -x -q -k "primary or secondary or guards or tiny or low_spp or mesh_bounded or megakernel_equals" 2>&1 | tail -3
echo "== 4-wide BVH vs binary (wavefront engine)"
timeout 900 python tools/ab.py build/rt_base.so build/rt_bvh4.so build/rt_bvh4tlas.so -- "--workload c2 --engine wavefront" "--workload c4 --engine wavefront" "--workload c5 --engine wavefront"
echo "== megakernel variants"
timeout 900 python tools/ab.py build/rt_base.so build/rt_regen8.so build/rt_pb8.so build/rt_pb8regen8.so build/rt_bvh4.so -- "--workload c1 --engine megakernel" "--workload c2 --engine megakernel" "--workload c3 --engine megakernel"
} > gpurun_out/r2_ab2.log 2>&1
cut -c1-200 gpurun_out/r2_ab2.log
echo "== default bench line (N=1)"
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_r2a_1gpu.json 2> gpurun_out/bench_r2a_1gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r2a_1gpu.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_r2a_ref.json 2> gpurun_out/bench_r2a_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_r2a_ref.json
echo "== steady-state ncu capture (k_trace / k_shade launches from the middle of a frame)"
BENCH="python bench.py --steps 1 --warmup 3 --spp 256 --no-cpu-baseline --no-e2e --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -s 130 -c 4 -o gpurun_out/prof_r2a -f $BENCH > gpurun_out/ncu_r2a.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/launches_r2a.csv $BENCH > gpurun_out/ncu_r2a_list.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/prof_r2a* gpurun_out/launches_r2a.csv
