#!/bin/bash
# Full-size bench + ncu evidence.  Usage: bash scripts/gpu_bench.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
$T 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_511_$TAG.json 2> gpurun_out/bench_511_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_511_$TAG.err
$T 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
# every launch with its device time (cold-cache, serialised): shares only
$T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --n 255 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
# the dominant kernel once, full set
$T 1200 ncu --set full --clock-control none --import-source on -k regex:cg_kernel -s 1 -c 1 -o gpurun_out/prof_cg_$TAG -f \
    python bench.py --n 255 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_511_$TAG.json; tail -2 gpurun_out/bench_511_$TAG.err; cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/ncu_full_$TAG.log
