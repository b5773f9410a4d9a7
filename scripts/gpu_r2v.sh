#!/bin/bash
# round 2, session 3: GPU suite with the node-parallel sweepers and the multi-implicit splitting; timings + ncu summaries
# of the kernels outside the headline step (GMRES, order-4 CG, reaction Newton)
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2v.log
grep -E "^E  |^FAILED|passed|failed|rc=" gpurun_out/pytest_gpu_r2v.log | tail -12
$T 300 python scripts/profile_extra.py > gpurun_out/extra_kernels_r2v.jsonl 2> gpurun_out/extra_kernels_r2v.err; cat gpurun_out/extra_kernels_r2v.jsonl; tail -3 gpurun_out/extra_kernels_r2v.err
F="--set full --clock-control none --import-source on -f"
$T 400 ncu $F -k regex:gmres_kernel -c 1 -o gpurun_out/prof_gmres_r02 python scripts/profile_extra.py gmres > gpurun_out/ncu_gmres.log 2>&1
$T 400 ncu $F -k regex:ho_cg_kernel -c 1 -o gpurun_out/prof_ho_cg_r02 python scripts/profile_extra.py ho_cg > gpurun_out/ncu_hocg.log 2>&1
$T 300 ncu $F -k regex:reaction_newton -c 1 -o gpurun_out/prof_reaction_r02 python scripts/profile_extra.py reaction > gpurun_out/ncu_react.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
