#!/bin/bash
# DRAM traffic of the CG kernel at the full 511^3 size (few metrics -> a single replay pass per launch).
TAG=${1:-r01b}
mkdir -p gpurun_out
timeout -k 10 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:cg_pipe_kernel -c 8 --csv --log-file gpurun_out/traffic_511_$TAG.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/traffic_511_$TAG.log 2>&1
tail -3 gpurun_out/traffic_511_$TAG.log; tail -6 gpurun_out/traffic_511_$TAG.csv
