#!/bin/bash
# 2-GPU diagnosis: where does the time outside the solver go at N=2?  (+ timeline sanity check on a tiny slab)
N=2
mkdir -p gpurun_out
T="timeout -k 10"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T 600 $RUN bench.py --gpus $N --steps 4 --warmup 3 --timeline --no-cpu-baseline --no-reference-controller > gpurun_out/bench_c3_2gpu_timeline_r2h.json 2> gpurun_out/bench_c3_2gpu_timeline_r2h.err; echo "rc=$?" >> gpurun_out/bench_c3_2gpu_timeline_r2h.err
$T 600 $RUN bench.py --gpus $N --n 127 --steps 4 --warmup 3 --timeline --no-cpu-baseline --no-reference-controller > gpurun_out/bench_c3_2gpu_n127_timeline_r2h.json 2> gpurun_out/bench_c3_2gpu_n127_timeline_r2h.err
$T 600 $RUN bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_2gpu_r2h.json 2> gpurun_out/bench_c3_2gpu_r2h.err; echo "rc=$?" >> gpurun_out/bench_c3_2gpu_r2h.err
$T 600 python bench.py --steps 4 --warmup 3 --timeline --no-cpu-baseline --no-reference-controller > gpurun_out/bench_c3_1gpu_timeline_r2h.json 2> gpurun_out/bench_c3_1gpu_timeline_r2h.err
python - <<'PY'
import json
for f in ["bench_c3_2gpu_timeline_r2h","bench_c3_2gpu_n127_timeline_r2h","bench_c3_2gpu_r2h","bench_c3_1gpu_timeline_r2h"]:
    try:
        d=[json.loads(l) for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        r=d["roofline"]
        print(f, "value %.4g e2e %.4g ms/step %.1f cg ms/launch %.2f share %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["ms_per_launch"],r["share_of_step"]))
        print("   ", d.get("timeline"))
    except Exception as e:
        print(f, "ERR", e)
PY
