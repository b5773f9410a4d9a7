#!/bin/bash
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2k.log; tail -2 gpurun_out/pytest_gpu_r2k.log
$T 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_r2k.json 2> gpurun_out/bench_c3_r2k.err; echo "rc=$?" >> gpurun_out/bench_c3_r2k.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_c3_r2k.json") if l.startswith("{")][-1]; r=d["roofline"]
print("config 3 value %.4g e2e %.4g ms/step %.1f frac %.3f traffic/alg %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"],r["traffic_over_algorithmic"]), d["check"]["status"], {k[:14]:round(v["frac_of_peak"],3) for k,v in d.get("other_kernels",{}).items()})
PY
