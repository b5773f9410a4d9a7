#!/bin/bash
# round 2, final code: what the driver runs at round end (GPU suite, smoke, default bench line)
mkdir -p gpurun_out
T="timeout -k 10"
$T 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2final.log
grep -E "^E  |^FAILED|passed|failed|rc=" gpurun_out/pytest_gpu_r2final.log | tail -8
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2final.log 2>&1; tail -2 gpurun_out/smoke_r2final.log
( time $T 900 python bench.py --gpus 1 --steps 8 --warmup 4 --cpu-budget 20 ) > gpurun_out/bench_c3_r2final.json 2> gpurun_out/bench_c3_r2final.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/bench_c3_r2final.json") if l.startswith("{")][-1]; r=d["roofline"]
print("value %.4g e2e %.4g ms/step %.1f frac %.3f traffic/alg %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["frac"],r.get("traffic_over_algorithmic")), d["check"]["status"], d["cpu_baseline"], {k[:14]:round(v["frac_of_peak"],3) for k,v in d.get("other_kernels",{}).items()}, d["clocks"], "launches", d.get("gpu_launches"))
PY
tail -4 gpurun_out/bench_c3_r2final.err
