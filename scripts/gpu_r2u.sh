#!/bin/bash
# round 2, session 3: GPU suite at HEAD + ncu summaries of the kernels added late in the round (GMRES, order 4/6/8)
mkdir -p gpurun_out
T="timeout -k 10"
$T 600 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2u.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu_r2u.log | tail -8
F="--set full --clock-control none --import-source on -f"
$T 300 ncu $F -k regex:gmres -c 1 -o gpurun_out/prof_gmres_r02 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gmres_against_oracle and 130" > gpurun_out/ncu_gmres.log 2>&1
$T 300 ncu $F -k regex:ho_cg_kernel -c 1 -o gpurun_out/prof_ho_cg_r02 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gmres_against_oracle and 127" > gpurun_out/ncu_hocg.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
