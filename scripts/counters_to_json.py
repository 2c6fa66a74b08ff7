#!/usr/bin/env python
"""gpurun_out/counters/*.csv (scripts/collect_counters.sh) -> profiles/ncu_kernel_counters.json: per kernel and launch size
the warp instructions and DRAM bytes of ONE launch in the steady state (bench.py reads it for roofline.traffic and for the
issue-slot roofline).

  python scripts/counters_to_json.py [label]     label = the directory under profiles/ the CSVs are kept in (default r02b_counters)

Entries are MERGED into the existing file, so a pass that re-collects only some launch sizes leaves the others alone.  A CSV
whose name starts with "robots_" holds every launch of one step of a bound robot group (scripts/run_robots.py): its kernels
are summed per kernel name and over the step (key robots_step@BxT, n_robots from the name)."""
import csv
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LABEL = sys.argv[1] if len(sys.argv) > 1 else "r02b_counters"
DEST = os.path.join(ROOT, "profiles", "ncu_kernel_counters.json")
merged = json.load(open(DEST)) if os.path.exists(DEST) else {}
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
    if name.startswith("robots_"):
        step = {"warp_inst": 0, "dram_bytes": 0, "ncu_duration_us": 0.0, "launches": 0, "per_kernel": {},
                "n_robots": int(re.search(r"robots_(\d+)_", name).group(1)), "source": f"profiles/{LABEL}/{name}.csv"}
        for (_, k), d in per.items():
            pk = step["per_kernel"].setdefault(k, {"warp_inst": 0, "dram_bytes": 0, "ncu_duration_us": 0.0, "launches": 0})
            for tgt in (pk, step):
                tgt["warp_inst"] += int(d.get("smsp__inst_executed.sum", 0))
                tgt["dram_bytes"] += int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
                tgt["ncu_duration_us"] += d.get("gpu__time_duration.sum") or 0.0
                tgt["launches"] += 1
        out[f"robots_step@{size}"] = step
        continue
    tag = "dense:" if "dense" in name else ""
    for (_, k), d in per.items():
        out[f"{tag}{k}@{size}"] = {"warp_inst": int(d.get("smsp__inst_executed.sum", 0)),
                                   "dram_bytes": int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)),
                                   "ncu_duration_us": d.get("gpu__time_duration.sum"),
                                   "source": f"profiles/{LABEL}/{name}.csv"}
merged.update(out)
json.dump(merged, open(DEST, "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
