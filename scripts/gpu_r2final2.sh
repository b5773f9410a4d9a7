#!/bin/bash
# round 2, final code (after the reference test-suite was added): GPU suite + smoke
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -m gpu -q --timeout=300 --durations=12 > gpurun_out/pytest_gpu_r2final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2final2.log
grep -E "^E  |^FAILED|passed|failed|rc=|s call" gpurun_out/pytest_gpu_r2final2.log | tail -22
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2final2.log 2>&1; tail -1 gpurun_out/smoke_r2final2.log
