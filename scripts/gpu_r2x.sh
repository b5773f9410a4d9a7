#!/bin/bash
# round 2, session 3: streaming passes of the GMRES / order-4..8 CG / reaction kernels as quad loops, stencils templated on
# the half width: GPU suite, timings, memcheck on the changed kernels, ncu re-captures
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2x.log
grep -E "^E  |^FAILED|passed|failed|rc=" gpurun_out/pytest_gpu_r2x.log | tail -12
$T 300 python scripts/profile_extra.py > gpurun_out/extra_kernels_r2x.jsonl 2> gpurun_out/extra_kernels_r2x.err; cat gpurun_out/extra_kernels_r2x.jsonl; tail -3 gpurun_out/extra_kernels_r2x.err
SEL="gmres_against_oracle or reaction_newton or spatial_accuracy or (test_operator and (o4 or o6 or o8 or advection or multi)) or (test_run and (gmres or multi or o4 or o8))"
$T 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/memcheck_r2x.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/memcheck_r2x.log | tail -3
F="--set full --clock-control none --import-source on -f"
$T 400 ncu $F -k regex:gmres_kernel -c 1 -o gpurun_out/prof_gmres_r02 python scripts/profile_extra.py gmres > gpurun_out/ncu_gmres.log 2>&1
$T 400 ncu $F -k regex:ho_cg_kernel -c 1 -o gpurun_out/prof_ho_cg_r02 python scripts/profile_extra.py ho_cg > gpurun_out/ncu_hocg.log 2>&1
$T 300 ncu $F -k regex:reaction_newton -c 1 -o gpurun_out/prof_reaction_r02 python scripts/profile_extra.py reaction > gpurun_out/ncu_react.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
