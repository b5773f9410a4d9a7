#!/bin/bash
# round 2, session 3, 4 GPUs: node-parallel sweepers over NCCL (one collocation node per GPU) - tests, then BASELINE config 3
# "parallel across the method"
mkdir -p gpurun_out
T="timeout -k 10"
nvidia-smi -L | head -8
$T 420 python -m pytest tests/test_node_parallel.py -m gpu -q --timeout=200 > gpurun_out/pytest_nodepar_4gpu_r2w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nodepar_4gpu_r2w.log
grep -E "^E  |^FAILED|passed|failed|rc=" gpurun_out/pytest_nodepar_4gpu_r2w.log | tail -8
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 scripts/bench_node_parallel.py --steps 4 --warmup 2 > gpurun_out/bench_nodepar_4gpu_r2w.json 2> gpurun_out/bench_nodepar_4gpu_r2w.err; echo "rc=$?" >> gpurun_out/bench_nodepar_4gpu_r2w.err
grep "^{" gpurun_out/bench_nodepar_4gpu_r2w.json | cut -c1-1500; tail -5 gpurun_out/bench_nodepar_4gpu_r2w.err
