#!/bin/bash
# round 2, session 3, 4 GPUs: nodes x slabs (2-D process grid) on the real kernels
mkdir -p gpurun_out
T="timeout -k 10"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513"
$T 240 $RUN tests/mgpu/grid_check.py 63 > gpurun_out/grid_check_r2z_4gpu_63.log 2>&1; echo "rc=$?" >> gpurun_out/grid_check_r2z_4gpu_63.log
grep -E "grid_check|rc=|Error|error" gpurun_out/grid_check_r2z_4gpu_63.log | tail -6
