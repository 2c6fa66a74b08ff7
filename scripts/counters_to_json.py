#!/usr/bin/env python
"""gpurun_out/counters/*.csv (scripts/collect_counters.sh) -> profiles/ncu_kernel_counters.json: per kernel and launch size
the warp instructions and DRAM bytes of ONE launch in the steady state (bench.py reads it for roofline.traffic and for the
issue-slot roofline)."""
import csv
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"_comment": "ncu --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                   "--clock-control none, one steady-state cycle (30 warm cycles before it) per launch size; "
                   "scripts/collect_counters.sh + scripts/counters_to_json.py; key = kernel@BxT"}
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "counters", "*.csv"))):
    name = os.path.basename(path)[:-4]
    m = re.search(r"(\d+)x(\d+)$", name)
    size = f"{m.group(1)}x{m.group(2)}"
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    if not rows:
        continue
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    idi = hdr.index("ID")
    per = {}
    for r in rows[1:]:
        k = re.sub(r"^void ", "", r[ki]).split("<")[0].split("(")[0].replace("mppi::", "")
        d = per.setdefault((r[idi], k), {})
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(unit, 1.0)
        d[r[mi]] = v * scale
    for (_, k), d in per.items():
        out[f"{k}@{size}"] = {"warp_inst": int(d.get("smsp__inst_executed.sum", 0)),
                              "dram_bytes": int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)),
                              "ncu_duration_us": d.get("gpu__time_duration.sum"),
                              "source": f"profiles/r02b_counters/{name}.csv"}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_kernel_counters.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
