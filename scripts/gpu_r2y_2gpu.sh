#!/bin/bash
# round 2, session 3, 2 GPUs: the slab path at the final code - parity check vs 1 GPU and the oracle, the 2-GPU test of the
# suite, slab bench with its answer check
mkdir -p gpurun_out
T="timeout -k 10"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T 300 $RUN tests/mgpu/slab_check.py 127 > gpurun_out/slab_check_r2y_2gpu_127.log 2>&1; echo "rc=$?" >> gpurun_out/slab_check_r2y_2gpu_127.log
grep -E "slab_check|rc=|Error|error" gpurun_out/slab_check_r2y_2gpu_127.log | tail -4
$T 400 python -m pytest tests/test_gpu_parity.py tests/test_node_parallel.py -m gpu -q --timeout=300 -k "two_gpus or equals_serial" > gpurun_out/pytest_r2y.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2y.log
tail -3 gpurun_out/pytest_r2y.log
$T 600 $RUN bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3_2gpu_r2y.json 2> gpurun_out/bench_c3_2gpu_r2y.err; echo "rc=$?" >> gpurun_out/bench_c3_2gpu_r2y.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_c3_2gpu_r2y.json") if l.startswith("{")][-1]
print("2 GPUs value %.4g e2e %.4g ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["check"]["status"], d["roofline"]["frac"], d["clocks"])
PY
tail -n 3 gpurun_out/bench_c3_2gpu_r2y.err
