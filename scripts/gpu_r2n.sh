#!/bin/bash
# A/B: 64x32 tiles, ONE CTA per SM with 16 consumer warps (default) vs 64x16 tiles, two CTAs per SM with 8 consumer warps each (-DSDCB200_SHORT_TILES)
mkdir -p gpurun_out
T="timeout -k 10"
B="python bench.py --no-cpu-baseline --no-reference-controller"
run() { tag=$1
  $T 900 python -m pytest tests -m gpu -q --timeout=300 -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log; tail -2 gpurun_out/pytest_$tag.log
  for c in 3 2 4; do $T 600 $B --config $c --steps 3 --warmup 2 --timeline > gpurun_out/bench_c${c}_$tag.json 2> gpurun_out/bench_c${c}_$tag.err; done
  python - <<PY
import json
for c in (3,2,4):
    try:
        d=[json.loads(l) for l in open("gpurun_out/bench_c%d_$tag.json"%c) if l.startswith("{")][-1]; r=d["roofline"]
        print("$tag config",c,"value %.4g ms/step %.1f frac %.3f ms/launch %.2f"%(d["value"],d["ms_per_step"],r["frac"],r["ms_per_launch"]), d.get("timeline",{}).get("launch_shape"), d["check"]["status"], {k[:14]:round(v["frac_of_peak"],3) for k,v in d.get("other_kernels",{}).items()})
    except Exception as e: print("$tag config",c,"ERR",e)
PY
}
run r2n_tall
SDCB200_NVCC_DEFS=-DSDCB200_SHORT_TILES python -m pysdc_b200.build --force > /dev/null 2>&1
run r2n_short
