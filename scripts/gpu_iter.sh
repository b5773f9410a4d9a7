#!/bin/bash
# Quick iteration on the GPU box: parity tests, then the headline bench.  Usage: bash scripts/gpu_iter.sh [tag] [bench args]
TAG=${1:-it}
shift
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
$T 900 python bench.py --steps 3 --warmup 3 "$@" > gpurun_out/bench_511_$TAG.json 2> gpurun_out/bench_511_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_511_$TAG.err
tail -15 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/bench_511_$TAG.json; tail -5 gpurun_out/bench_511_$TAG.err
