#!/bin/bash
# round 2: what the driver runs at round end, on the final code
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2l.log; tail -2 gpurun_out/pytest_gpu_r2l.log
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2l.log 2>&1; tail -1 gpurun_out/smoke_r2l.log
( time $T 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_driverlike_r2l.json 2> gpurun_out/bench_driverlike_r2l.err
( time $T 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_reference_driverlike_r2l.json 2> gpurun_out/bench_reference_driverlike_r2l.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_driverlike_r2l.json") if l.startswith("{")][-1]; r=d["roofline"]
print("value %.4g e2e %.4g ms/step %.1f frac %.3f traffic/alg %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"],r["traffic_over_algorithmic"]), d["check"]["status"], d["cpu_baseline"], {k[:14]:round(v["frac_of_peak"],3) for k,v in d.get("other_kernels",{}).items()}, d["clocks"])
d=[json.loads(l) for l in open("gpurun_out/bench_reference_driverlike_r2l.json") if l.startswith("{")][-1]
print("reference arm value %.4g steps %d ms/step %.0f"%(d["value"], d["steps"], d["ms_per_step"]), d["config"], d["cpu_baseline"]["sample"][:200])
PY
tail -4 gpurun_out/bench_driverlike_r2l.err gpurun_out/bench_reference_driverlike_r2l.err
