#!/bin/bash
# round 2, second session: answer records, bench lines for configs 3 / 2 / 4, the reference arm as the driver runs it
mkdir -p gpurun_out
T="timeout -k 10"
for c in 3 2 4; do
  $T 600 python bench.py --config $c --steps 1 --warmup 1 --write-record --no-cpu-baseline > gpurun_out/record_c$c.json 2> gpurun_out/record_c$c.err; echo "rc=$?" >> gpurun_out/record_c$c.err
done
cp tests/golden/bench_record_*.json gpurun_out/
$T 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3_r2b.json 2> gpurun_out/bench_c3_r2b.err; echo "rc=$?" >> gpurun_out/bench_c3_r2b.err
$T 600 python bench.py --config 2 --steps 5 --warmup 3 > gpurun_out/bench_c2_r2b.json 2> gpurun_out/bench_c2_r2b.err; echo "rc=$?" >> gpurun_out/bench_c2_r2b.err
$T 600 python bench.py --config 4 --steps 3 --warmup 2 > gpurun_out/bench_c4_r2b.json 2> gpurun_out/bench_c4_r2b.err; echo "rc=$?" >> gpurun_out/bench_c4_r2b.err
( time $T 900 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_ref_r2b.json 2> gpurun_out/bench_ref_r2b.err
for f in record_c3 record_c2 record_c4 bench_c3_r2b bench_c2_r2b bench_c4_r2b bench_ref_r2b; do echo "== $f"; cut -c1-3000 gpurun_out/$f.json; tail -4 gpurun_out/$f.err; done
