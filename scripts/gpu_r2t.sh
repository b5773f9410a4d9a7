#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2t.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu_r2t.log | tail -12
