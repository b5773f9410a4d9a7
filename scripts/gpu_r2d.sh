#!/bin/bash
# round 2, fourth session: GPU suite (higher-order stencils, new z-chunk rule) + ncu evidence for every kernel
mkdir -p gpurun_out
T="timeout -k 10"
B="python bench.py --no-cpu-baseline --no-reference-controller"
$T 1200 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2d.log
M="--metrics gpu__time_duration.sum --clock-control none --csv"
$T 600 ncu $M -c 200 --log-file gpurun_out/launches_r02_config3.csv $B --steps 1 --warmup 1 > gpurun_out/ncu_l3.log 2>&1
$T 600 ncu $M -c 400 --log-file gpurun_out/launches_r02_config2.csv $B --config 2 --steps 1 --warmup 1 > gpurun_out/ncu_l2.log 2>&1
$T 600 ncu $M -c 200 --log-file gpurun_out/launches_r02_config4.csv $B --config 4 --steps 1 --warmup 1 > gpurun_out/ncu_l4.log 2>&1
D="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv"
$T 900 ncu $D -k regex:cg_pipe_kernel -c 4 --log-file gpurun_out/traffic_c3.csv $B --steps 1 --warmup 0 > gpurun_out/ncu_t3.log 2>&1
$T 600 ncu $D -k regex:cg_pipe_kernel -c 12 --log-file gpurun_out/traffic_c2.csv $B --config 2 --steps 1 --warmup 0 > gpurun_out/ncu_t2.log 2>&1
$T 600 ncu $D -k regex:newton_pipe_kernel -c 3 --log-file gpurun_out/traffic_c4.csv $B --config 4 --steps 1 --warmup 0 > gpurun_out/ncu_t4.log 2>&1
$T 600 ncu $D -k "regex:eval_pipe|colloc" -c 14 --log-file gpurun_out/traffic_stream.csv $B --steps 1 --warmup 0 > gpurun_out/ncu_ts.log 2>&1
F="--set full --clock-control none --import-source on -f"
$T 900 ncu $F -k regex:newton_pipe_kernel -c 1 -o gpurun_out/prof_newton_pipe_r02 $B --config 4 --n 1024 --steps 1 --warmup 0 > gpurun_out/ncu_f4.log 2>&1
$T 900 ncu $F -k regex:cg_pipe_kernel -s 4 -c 1 -o gpurun_out/prof_cg_pipe2d_r02 $B --config 2 --steps 1 --warmup 0 > gpurun_out/ncu_f2.log 2>&1
$T 900 ncu $F -k regex:eval_pipe_kernel -s 1 -c 1 -o gpurun_out/prof_eval_pipe_r02 $B --n 255 --steps 1 --warmup 0 > gpurun_out/ncu_f3e.log 2>&1
$T 900 ncu $F -k regex:colloc_sweep_kernel -c 1 -o gpurun_out/prof_colloc_sweep_r02 $B --n 255 --steps 1 --warmup 0 > gpurun_out/ncu_f3c.log 2>&1
$T 900 ncu $F -k regex:colloc_residual_kernel -c 1 -o gpurun_out/prof_colloc_residual_r02 $B --n 255 --steps 1 --warmup 0 > gpurun_out/ncu_f3r.log 2>&1
$T 900 ncu $F -k regex:cg_pipe_kernel -s 1 -c 1 -o gpurun_out/prof_cg_pipe3d_r02 $B --n 255 --steps 1 --warmup 0 > gpurun_out/ncu_f3.log 2>&1
tail -15 gpurun_out/pytest_gpu_r2d.log; ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv; tail -3 gpurun_out/ncu_f4.log gpurun_out/ncu_t3.log
