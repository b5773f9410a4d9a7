#!/bin/bash
# round 2, 8-GPU session: slab check, slab bench (answer check + timeline), PFASST config 5 over NCCL (check + bench line)
N=${1:-8}
mkdir -p gpurun_out
T="timeout -k 10"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T 300 $RUN tests/mgpu/slab_check.py 127 > gpurun_out/slab_check_r2g_${N}gpu_127.log 2>&1; echo "rc=$?" >> gpurun_out/slab_check_r2g_${N}gpu_127.log
grep -E "slab_check|rc=|Error|error" gpurun_out/slab_check_r2g_${N}gpu_127.log | tail -4
$T 600 $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_${N}gpu_r2g.json 2> gpurun_out/bench_c3_${N}gpu_r2g.err; echo "rc=$?" >> gpurun_out/bench_c3_${N}gpu_r2g.err
cut -c1-3000 gpurun_out/bench_c3_${N}gpu_r2g.json; tail -n 3 gpurun_out/bench_c3_${N}gpu_r2g.err
$T 600 $RUN bench.py --gpus $N --steps 3 --warmup 2 --timeline --no-cpu-baseline --no-reference-controller > gpurun_out/bench_c3_${N}gpu_timeline_r2g.json 2> gpurun_out/bench_c3_${N}gpu_timeline_r2g.err; echo "rc=$?" >> gpurun_out/bench_c3_${N}gpu_timeline_r2g.err
cut -c1-3600 gpurun_out/bench_c3_${N}gpu_timeline_r2g.json; tail -n 3 gpurun_out/bench_c3_${N}gpu_timeline_r2g.err
if [ "$N" = "8" ]; then
  $T 600 $RUN tests/mgpu/pfasst_check.py pfasst_config5_1023_p8 nccl > gpurun_out/pfasst_config5_r2g.log 2>&1; echo "rc=$?" >> gpurun_out/pfasst_config5_r2g.log
  grep -E "pfasst_check|rc=|Error" gpurun_out/pfasst_config5_r2g.log | tail -3
  $T 600 $RUN tests/mgpu/pfasst_check.py pfasst_step8A_heat1d nccl > gpurun_out/pfasst_step8A_r2g.log 2>&1; echo "rc=$?" >> gpurun_out/pfasst_step8A_r2g.log
  grep -E "pfasst_check|rc=|Error" gpurun_out/pfasst_step8A_r2g.log | tail -3
  $T 600 $RUN bench.py --gpus 8 --config 5 --steps 5 --warmup 2 --write-record --no-cpu-baseline > gpurun_out/bench_c5_8gpu_r2g.json 2> gpurun_out/bench_c5_8gpu_r2g.err; echo "rc=$?" >> gpurun_out/bench_c5_8gpu_r2g.err
  cp tests/golden/bench_record_config5_n1023_p8.json gpurun_out/ 2>/dev/null
  cut -c1-3000 gpurun_out/bench_c5_8gpu_r2g.json; tail -n 3 gpurun_out/bench_c5_8gpu_r2g.err
fi
