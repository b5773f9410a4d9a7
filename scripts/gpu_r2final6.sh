#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests/test_reference_suite.py -m gpu -q --timeout=150 -k "cpu_vs_gpu" > gpurun_out/pytest_r2final6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2final6.log
grep -E "^E  |^FAILED|passed|failed|rc=" gpurun_out/pytest_r2final6.log | tail -5
