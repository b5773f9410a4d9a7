#!/bin/bash
# round 2, final single-GPU session: GPU suite, compute-sanitizer on the new kernels, ncu re-captures at the final code,
# driver-like bench lines
mkdir -p gpurun_out
T="timeout -k 10"
B="python bench.py --no-cpu-baseline --no-reference-controller"
$T 1200 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2j.log
tail -3 gpurun_out/pytest_gpu_r2j.log
# sanitizer: periodic / variable-diagonal / batched Newton / pipelined eval_f / higher-order kernels at small sizes
SEL="stencil_and_cg_against_oracle or batched_newton or batched_solve or cg_edge or operator or streaming or sweep_combinations"
$T 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/memcheck_r2j.log 2>&1
$T 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/synccheck_r2j.log 2>&1
$T 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batched_newton or cg_edge or (stencil_and_cg_against_oracle and (2-4- or 2-66- or 3-4- or 3-34- or 2-65- or 3-33-))" > gpurun_out/racecheck_r2j.log 2>&1
for f in memcheck synccheck racecheck; do echo "## $f"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${f}_r2j.log | tail -3; done
# ncu at the final code
D="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv"
$T 900 ncu $D -k regex:cg_pipe_kernel -c 4 --log-file gpurun_out/traffic_c3.csv $B --steps 1 --warmup 0 > gpurun_out/ncu_t3.log 2>&1
$T 600 ncu $D -k regex:cg_pipe_kernel -c 12 --log-file gpurun_out/traffic_c2.csv $B --config 2 --steps 1 --warmup 0 > gpurun_out/ncu_t2.log 2>&1
$T 600 ncu $D -k regex:newton_pipe_kernel -c 3 --log-file gpurun_out/traffic_c4.csv $B --config 4 --steps 1 --warmup 0 > gpurun_out/ncu_t4.log 2>&1
$T 600 ncu $D -k "regex:eval_pipe|colloc" -c 14 --log-file gpurun_out/traffic_stream.csv $B --steps 1 --warmup 0 > gpurun_out/ncu_ts.log 2>&1
F="--set full --clock-control none --import-source on -f"
$T 900 ncu $F -k regex:newton_pipe_kernel -c 1 -o gpurun_out/prof_newton_pipe_r02 $B --config 4 --steps 1 --warmup 0 > gpurun_out/ncu_f4.log 2>&1
$T 900 ncu $F -k regex:cg_pipe_kernel -s 4 -c 1 -o gpurun_out/prof_cg_pipe2d_r02 $B --config 2 --steps 1 --warmup 0 > gpurun_out/ncu_f2.log 2>&1
$T 900 ncu $F -k regex:eval_pipe_kernel -s 1 -c 1 -o gpurun_out/prof_eval_pipe_r02 $B --n 255 --steps 1 --warmup 0 > gpurun_out/ncu_f3e.log 2>&1
$T 900 ncu $F -k regex:colloc_residual_kernel -c 1 -o gpurun_out/prof_colloc_residual_r02 $B --n 255 --steps 1 --warmup 0 > gpurun_out/ncu_f3r.log 2>&1
$T 900 ncu $F -k regex:colloc_sweep_kernel -c 1 -o gpurun_out/prof_colloc_sweep_r02 $B --n 255 --steps 1 --warmup 0 > gpurun_out/ncu_f3c.log 2>&1
$T 900 ncu $F -k regex:cg_pipe_kernel -s 1 -c 1 -o gpurun_out/prof_cg_pipe3d_r02 $B --n 255 --steps 1 --warmup 0 > gpurun_out/ncu_f3.log 2>&1
M="--metrics gpu__time_duration.sum --clock-control none --csv"
$T 600 ncu $M -c 200 --log-file gpurun_out/launches_r02c3.csv $B --steps 1 --warmup 1 > gpurun_out/ncu_l3.log 2>&1
$T 600 ncu $M -c 400 --log-file gpurun_out/launches_r02c2.csv $B --config 2 --steps 1 --warmup 1 > gpurun_out/ncu_l2.log 2>&1
$T 600 ncu $M -c 200 --log-file gpurun_out/launches_r02c4.csv $B --config 4 --steps 1 --warmup 1 > gpurun_out/ncu_l4.log 2>&1
# bench lines (config 3 as the driver runs it, but fewer steps; configs 2 / 4 with their cpu baselines)
$T 900 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_c3_r2j.json 2> gpurun_out/bench_c3_r2j.err; echo "rc=$?" >> gpurun_out/bench_c3_r2j.err
$T 600 python bench.py --config 2 --steps 5 --warmup 3 --cpu-budget 40 > gpurun_out/bench_c2_r2j.json 2> gpurun_out/bench_c2_r2j.err; echo "rc=$?" >> gpurun_out/bench_c2_r2j.err
$T 600 python bench.py --config 4 --steps 3 --warmup 2 --cpu-budget 40 > gpurun_out/bench_c4_r2j.json 2> gpurun_out/bench_c4_r2j.err; echo "rc=$?" >> gpurun_out/bench_c4_r2j.err
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2j.log 2>&1; tail -2 gpurun_out/smoke_r2j.log
python - <<'PY'
import json
for c in (3,2,4):
    try:
        d=[json.loads(l) for l in open("gpurun_out/bench_c%d_r2j.json"%c) if l.startswith("{")][-1]; r=d["roofline"]
        print("config",c,"value %.4g e2e %.4g ms/step %.1f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"]), d["check"]["status"], d.get("cpu_baseline",{}).get("value"), {k[:14]:round(v["frac_of_peak"],3) for k,v in d.get("other_kernels",{}).items()})
    except Exception as e: print("config",c,"ERR",e)
PY
