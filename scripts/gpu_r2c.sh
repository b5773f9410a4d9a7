#!/bin/bash
# round 2, third session: new pipelined kernels (periodic wrap, variable diagonal, batched Newton, eval_f): GPU suite + benches
mkdir -p gpurun_out
T="timeout -k 10"
$T 1500 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2c.log
$T 1200 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2c_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2c_all.log
for c in 3 2 4; do
  $T 600 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c${c}_r2c.json 2> gpurun_out/bench_c${c}_r2c.err; echo "rc=$?" >> gpurun_out/bench_c${c}_r2c.err
done
tail -25 gpurun_out/pytest_gpu_r2c.log; tail -12 gpurun_out/pytest_gpu_r2c_all.log
for c in 3 2 4; do echo "== c$c"; cut -c1-2600 gpurun_out/bench_c${c}_r2c.json; tail -4 gpurun_out/bench_c${c}_r2c.err; done
