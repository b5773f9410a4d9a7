#!/bin/bash
# scaling points at the final code: bench.py --gpus N with the timeline
N=${1:-2}
mkdir -p gpurun_out
T="timeout -k 10"
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T 600 $RUN bench.py --gpus $N --steps 5 --warmup 3 --timeline --no-cpu-baseline > gpurun_out/bench_c3_${N}gpu_r2m.json 2> gpurun_out/bench_c3_${N}gpu_r2m.err; echo "rc=$?" >> gpurun_out/bench_c3_${N}gpu_r2m.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_c3_${N}gpu_r2m.json") if l.startswith("{")][-1]; r=d["roofline"]
print("N=$N value %.4g e2e %.4g ms/step %.1f cg ms/launch %.2f share %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["ms_per_launch"],r["share_of_step"]), d["check"]["status"])
print("   ", d.get("timeline"))
PY
tail -n 3 gpurun_out/bench_c3_${N}gpu_r2m.err
