#!/bin/bash
# PFASST on 8 GPUs (one time slice per GPU, NCCL send/recv): the reference's fixtures incl. BASELINE config 5 at full size.
TAG=${1:-pf8}
shift
NAMES=${@:-pfasst_step8A_heat1d pfasst_heat2d_imex_127_p8 pfasst_config5_1023_p8}
mkdir -p gpurun_out
T="timeout -k 10"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
for name in $NAMES; do
  $T 600 $RUN tests/mgpu/pfasst_check.py $name nccl > gpurun_out/pfasst_${TAG}_$name.log 2>&1; echo "rc=$?" >> gpurun_out/pfasst_${TAG}_$name.log
  grep -E "pfasst_check|rc=|Error" gpurun_out/pfasst_${TAG}_$name.log | tail -4
done
