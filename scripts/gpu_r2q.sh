#!/bin/bash
# eval_f pipeline depth experiment, shallow side: 3 / 4 / 5 stages
mkdir -p gpurun_out
for v in 3 4 5; do
  SDCB200_NVCC_DEFS="-DSDCB200_EVAL_STAGES=$v" python -m pysdc_b200.build --force > /dev/null 2>&1
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-controller > gpurun_out/bench_evalst_$v.json 2>/dev/null
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_evalst_$v.json") if l.startswith("{")][-1]
print("eval stages $v:", {k[:14]:round(x["frac_of_peak"],3) for k,x in d["other_kernels"].items()}, "value %.4g"%d["value"])
PY
done
