#!/bin/bash
# ncu evidence for profiles/: launch list (shares) + one full capture of the dominant kernel.  Usage: gpu_profile.sh tag [kernel-regex] [n]
TAG=${1:-r01}
KRE=${2:-cg_pipe_kernel}
N=${3:-255}
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --n $N --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
$T 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -o gpurun_out/prof_cg_$TAG -f \
    python bench.py --n $N --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_launches_$TAG.log gpurun_out/ncu_full_$TAG.log
