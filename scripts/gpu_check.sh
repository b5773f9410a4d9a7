#!/bin/bash
# One GPU-box session: staged parity tests, smoke, short bench.  Logs go to gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
T="timeout -k 10"
$T 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=120 -k "streaming or datatype or device" > gpurun_out/pytest_stage1.log 2>&1
echo "stage1 rc=$?" >> gpurun_out/pytest_stage1.log
$T 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=120 -k "stencil_and_cg or batched or edge" > gpurun_out/pytest_stage2.log 2>&1
echo "stage2 rc=$?" >> gpurun_out/pytest_stage2.log
$T 1500 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu.log 2>&1
echo "all rc=$?" >> gpurun_out/pytest_gpu.log
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
for n in 127 255; do
  $T 600 python bench.py --n $n --steps 2 --warmup 1 --ref-n 31 > gpurun_out/bench_n$n.log 2>&1
  echo "bench $n rc=$?" >> gpurun_out/bench_n$n.log
done
tail -5 gpurun_out/pytest_stage1.log gpurun_out/pytest_stage2.log gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_n127.log gpurun_out/bench_n255.log
