#!/bin/bash
# round 2, fifth session: suite + benches after the eval pipeline / collocation / z-chunk changes, timeline, config 5 dry run
mkdir -p gpurun_out
T="timeout -k 10"
B="python bench.py --no-cpu-baseline"
$T 1200 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2e.log
$T 600 $B --steps 3 --warmup 3 > gpurun_out/bench_c3_r2e.json 2> gpurun_out/bench_c3_r2e.err; echo "rc=$?" >> gpurun_out/bench_c3_r2e.err
$T 600 $B --steps 2 --warmup 1 --timeline --no-reference-controller > gpurun_out/bench_c3_timeline_r2e.json 2> gpurun_out/bench_c3_timeline_r2e.err; echo "rc=$?" >> gpurun_out/bench_c3_timeline_r2e.err
$T 600 $B --config 2 --steps 3 --warmup 2 > gpurun_out/bench_c2_r2e.json 2> gpurun_out/bench_c2_r2e.err; echo "rc=$?" >> gpurun_out/bench_c2_r2e.err
$T 600 $B --config 4 --steps 3 --warmup 2 > gpurun_out/bench_c4_r2e.json 2> gpurun_out/bench_c4_r2e.err; echo "rc=$?" >> gpurun_out/bench_c4_r2e.err
$T 600 $B --config 5 --n 255 --steps 2 --warmup 1 > gpurun_out/bench_c5_dry_r2e.json 2> gpurun_out/bench_c5_dry_r2e.err; echo "rc=$?" >> gpurun_out/bench_c5_dry_r2e.err
$T 600 $B --precond --steps 2 --warmup 1 --no-reference-controller > gpurun_out/bench_c3_precond_r2e.json 2> gpurun_out/bench_c3_precond_r2e.err; echo "rc=$?" >> gpurun_out/bench_c3_precond_r2e.err
tail -8 gpurun_out/pytest_gpu_r2e.log
for f in bench_c3_r2e bench_c3_timeline_r2e bench_c2_r2e bench_c4_r2e bench_c5_dry_r2e bench_c3_precond_r2e; do echo "== $f"; cut -c1-3200 gpurun_out/$f.json; tail -n 4 gpurun_out/$f.err; done
