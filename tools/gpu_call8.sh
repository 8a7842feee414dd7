#!/bin/bash
# round 2, GPU call 8 (8 GPUs): where a strong-scaled C4 frame spends its time - render / collective / resolve per rank,
# reduce against the exchange of owned tiles, the collectives alone
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/mgpu_diag.py --frames 12 > gpurun_out/r2_mgpu_diag8.log 2> gpurun_out/r2_mgpu_diag8.err; echo "diag rc=$?"
grep -v "^NCCL version" gpurun_out/r2_mgpu_diag8.log
tail -5 gpurun_out/r2_mgpu_diag8.err
