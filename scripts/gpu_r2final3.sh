#!/bin/bash
# round 2, final code: the reference's tests on the CUDA kernels (incl. test_check_convergence with the CG-to-attainable-
# accuracy 'direct' solves) - then the whole GPU suite once more
mkdir -p gpurun_out
T="timeout -k 10"
$T 600 python -m pytest tests/test_reference_suite.py -m gpu -q --timeout=200 --durations=6 > gpurun_out/pytest_refsuite_r2final3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_refsuite_r2final3.log
grep -E "^E  |^FAILED|passed|failed|rc=|s call" gpurun_out/pytest_refsuite_r2final3.log | tail -12
