#!/bin/bash
# round 2, 2-GPU session: slab parity checks, the 2-GPU test of the suite, slab bench with answer check + timeline
N=${1:-2}
mkdir -p gpurun_out
T="timeout -k 10"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for n in 63 127; do
  $T 300 $RUN tests/mgpu/slab_check.py $n > gpurun_out/slab_check_r2f_${N}gpu_$n.log 2>&1; echo "rc=$?" >> gpurun_out/slab_check_r2f_${N}gpu_$n.log
  grep -E "slab_check|rc=|Error|error" gpurun_out/slab_check_r2f_${N}gpu_$n.log | tail -4
done
$T 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -k "two_gpus or pfasst_time_slices or streaming or sweep_combinations" > gpurun_out/pytest_r2f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2f.log
tail -5 gpurun_out/pytest_r2f.log
$T 900 $RUN bench.py --gpus $N --steps 3 --warmup 2 --timeline --no-cpu-baseline > gpurun_out/bench_c3_${N}gpu_r2f.json 2> gpurun_out/bench_c3_${N}gpu_r2f.err; echo "rc=$?" >> gpurun_out/bench_c3_${N}gpu_r2f.err
cut -c1-3500 gpurun_out/bench_c3_${N}gpu_r2f.json; tail -n 5 gpurun_out/bench_c3_${N}gpu_r2f.err
$T 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-controller > gpurun_out/bench_c3_1gpu_r2f.json 2> gpurun_out/bench_c3_1gpu_r2f.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c3_1gpu_r2f.json").read().strip().splitlines()[-1])
print("1 GPU value %.4g"%d["value"], {k[:30]: round(v["frac_of_peak"],3) for k,v in d["other_kernels"].items()})
PY
