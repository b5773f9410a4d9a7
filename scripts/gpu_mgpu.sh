#!/bin/bash
# Multi-GPU session: slab parity checks, then the slab-sharded headline bench.  Usage: gpu_mgpu.sh NGPUS tag
N=${1:-2}
TAG=${2:-mg}
mkdir -p gpurun_out
T="timeout -k 10"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
for n in 31 63 127; do
  $T 300 $RUN tests/mgpu/slab_check.py $n > gpurun_out/slab_check_${TAG}_$n.log 2>&1; echo "rc=$?" >> gpurun_out/slab_check_${TAG}_$n.log
  grep -E "slab_check|rc=|Error|error" gpurun_out/slab_check_${TAG}_$n.log | tail -5
done
$T 900 $RUN bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_511_${TAG}.json 2> gpurun_out/bench_511_${TAG}.err; echo "bench rc=$?" >> gpurun_out/bench_511_${TAG}.err
cat gpurun_out/bench_511_${TAG}.json; tail -5 gpurun_out/bench_511_${TAG}.err
