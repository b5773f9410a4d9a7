#!/usr/bin/env python
"""ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file F) -> the
per-launch DRAM traffic records bench.py reads (profiles/*_traffic_*.json).

    python scripts/traffic_json.py gpurun_out/traffic_c3.csv profiles/cg_traffic_511.json "note ..." [kernel-substring]
"""
import csv
import json
import sys

src, dst, note = sys.argv[1:4]
want = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(src)))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ix = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
         "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
per = {}
for r in rows[h + 1:]:
    if len(r) <= ix["Metric Value"] or (want and want not in r[ix["Kernel Name"]]):
        continue
    d = per.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]].split("(")[0][-60:]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * scale[r[ix["Metric Unit"]]]
launches = []
for i, (k, d) in enumerate(sorted(per.items())):
    rd, wr, ms = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"], d["gpu__time_duration.sum"]
    launches.append(dict(launch=i, kernel=d["kernel"], dram_read=rd, dram_write=wr, traffic=rd + wr, ms=ms,
                         gbs=(rd + wr) / (ms * 1e-3) / 1e9))
json.dump(dict(note=note, launches=launches), open(dst, "w"), indent=1)
print(dst, [(l["kernel"][-24:], round(l["traffic"] / 1e9, 2), round(l["gbs"])) for l in launches])
