#!/bin/bash
# eval_f pipeline depth experiment: 5 stages with the tile-only slot (default) vs 6 / 7 / 8 stages without it
mkdir -p gpurun_out
for v in default 6 7 8; do
  if [ "$v" = "default" ]; then python -m pysdc_b200.build --force > /dev/null 2>&1; else SDCB200_NVCC_DEFS="-DSDCB200_EVAL_SLIM -DSDCB200_SLIM_STAGES=$v" python -m pysdc_b200.build --force > /dev/null 2>&1; fi
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-controller > gpurun_out/bench_eval_$v.json 2>/dev/null
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_eval_$v.json") if l.startswith("{")][-1]
print("eval stages $v:", {k[:14]:round(x["frac_of_peak"],3) for k,x in d["other_kernels"].items()}, "value %.4g"%d["value"])
PY
done
