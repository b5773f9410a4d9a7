#!/bin/bash
# round 2, first session: whole GPU suite (reference controller on the CUDA kernels included), configs 2/4, short bench
mkdir -p gpurun_out
T="timeout -k 10"
$T 1500 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2a.log
$T 600 python scripts/bench_configs.py --steps 2 > gpurun_out/bench_configs_r2a.jsonl 2> gpurun_out/bench_configs_r2a.err
$T 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_511_r2a.json 2> gpurun_out/bench_511_r2a.err; echo "bench rc=$?" >> gpurun_out/bench_511_r2a.err
tail -30 gpurun_out/pytest_gpu_r2a.log; tail -3 gpurun_out/smoke_r2a.log; cat gpurun_out/bench_configs_r2a.jsonl; tail -3 gpurun_out/bench_configs_r2a.err; cat gpurun_out/bench_511_r2a.json; tail -3 gpurun_out/bench_511_r2a.err
