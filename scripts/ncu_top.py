#!/usr/bin/env python
"""Print the key metrics / top utilisations / stall reasons of an .ncu-rep (first kernel).  Usage: ncu_top.py file.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
d = dict(zip(hdr, zip(vals, units)))
for k in ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
          "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
          "launch__block_size", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum"]:
    if k in d:
        print(f"{k} = {d[k][0]} {d[k][1]}")
print("-- stalls (warps per issue-active) --")
st = []
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            st.append((float(d[h][0].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
for v, h in sorted(st, reverse=True)[:10]:
    print(f"{v:8.3f} {h}")
