#!/bin/bash
# round 2, GPU call 1: engine tests, engine A/B on all five configs, k_trace and k_path knobs, then the whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2_call1_gpu.txt
timeout 600 python -m pytest tests/test_gpu_engines.py -x -q > gpurun_out/r2_engines_test.log 2>&1; echo "engines test rc=$?" | tee -a gpurun_out/r2_call1.log
{
echo "== engines, base library"
for w in c1 c2 c3 c4 c5; do
  timeout 300 python tools/ab.py build/rt_base.so -- "--workload $w --engine wavefront" "--workload $w --engine megakernel"
done
echo "== k_trace knobs (wavefront engine)"
timeout 900 python tools/ab.py build/rt_old.so build/rt_notlas.so build/rt_nopf.so build/rt_tlas512.so -- "--workload c4 --engine wavefront" "--workload c5 --engine wavefront"
echo "== k_path knobs (megakernel engine)"
timeout 1200 python tools/ab.py build/rt_pb4.so build/rt_pb5.so build/rt_pb8.so build/rt_regen8.so build/rt_regen32.so -- "--workload c1 --engine megakernel" "--workload c3 --engine megakernel" "--workload c4 --engine megakernel"
} > gpurun_out/r2_ab1.log 2>&1
tail -50 gpurun_out/r2_ab1.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu1.log 2>&1; echo "gpu suite rc=$?" | tee -a gpurun_out/r2_call1.log
tail -15 gpurun_out/r2_pytest_gpu1.log
