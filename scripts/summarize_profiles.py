#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> [kernel-regex-for-the-full-capture]
"""
import collections
import csv
import os
import subprocess
import sys

tag = sys.argv[1]
out_dir = "profiles"
os.makedirs(out_dir, exist_ok=True)

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
           "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
           "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct"]

launches = f"gpurun_out/launches_{tag}.csv"
if os.path.exists(launches):
    rows = list(csv.reader(open(launches)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv, mn, mu = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[mu]]
        a = agg.setdefault(r[kn].split("(")[0][-70:], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(f"{out_dir}/{tag}_launch_list.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold cache, serialised: compare SHARES)\n")
        f.write(f"# source: {launches}; total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        f.write("kernel, launches, total_ms, share\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k}, {c}, {t:.3f}, {t / tot:.4f}\n")
    print(open(f"{out_dir}/{tag}_launch_list.txt").read())

for rep in sorted(f for f in os.listdir("gpurun_out") if f.endswith(f"_{tag}.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name = rep[:-len(".ncu-rep")]
    with open(f"{out_dir}/{name}_summary.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; report {rep} (kept in gpurun_out/, not tracked)\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"{m} = {r[i]} {units[i]}\n")
    print(open(f"{out_dir}/{name}_summary.txt").read())
