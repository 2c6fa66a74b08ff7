#!/usr/bin/env python
"""Per-source-line summary of an ncu report: instructions executed and stall samples per CUDA line.

  python scripts/ncu_by_line.py gpurun_out/prof.ncu-rep rollout_score_kernel [--top 40] [--launch 0]

ncu's CSV source page is SASS-only, so the SASS rows are matched (by instruction offset) against
`nvdisasm -g` of the cubin inside mpcholonavigation_b200/libmppi_b200.so, which carries the -lineinfo
line markers.  The library must be the build that was profiled.
"""
import argparse
import csv
import glob
import io
import os
import re
import subprocess
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mpcholonavigation_b200", "libmppi_b200.so")


def disasm_lines(kernel):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, check=True, capture_output=True)
        cubin = glob.glob(os.path.join(d, "*.cubin"))[0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    off2line, cur, active = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return off2line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0, help="index among the launches of that kernel in the report")
    ap.add_argument("--sort", default="samples", choices=["samples", "inst"])
    ap.add_argument("--mangled", default=None, help="substring of the mangled name (templates), default = kernel")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the CSV holds one block per profiled launch: "Kernel Name", header row, rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    blocks = [b for b in blocks if a.kernel in b["name"]]
    blk = blocks[a.launch]
    hdr = blk["hdr"]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ithr = hdr.index("Thread Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    base = int(blk["rows"][0][ia], 16)
    off2line = disasm_lines(a.mangled or a.kernel)
    per_line = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot_i = tot_s = tot_t = 0
    for r in blk["rows"]:
        off = int(r[ia], 16) - base
        line, _ = off2line.get(off, (("?", 0), ""))
        n, s, t = int(r[ii] or 0), int(r[isamp] or 0), int(r[ithr] or 0)
        e = per_line[line]
        e[0] += n; e[1] += s; e[2] += t
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                e[3][hdr[c]] += v
        tot_i += n; tot_s += s; tot_t += t
    src = {}
    for f in glob.glob(os.path.join(ROOT, "mpcholonavigation_b200", "csrc", "*")) + glob.glob(os.path.join(ROOT, "include", "*")):
        try:
            src[os.path.basename(f)] = open(f).read().splitlines()
        except Exception:
            pass
    print(f"kernel {blk['name'][:70]}: {tot_i} warp instructions, {tot_t} thread instructions, {tot_s} stall samples")
    print("  inst%  samp%  file:line  top stalls | source")
    col = 1 if a.sort == "samples" else 0
    for line, e in sorted(per_line.items(), key=lambda kv: -kv[1][col])[: a.top]:
        f, n = line if line else ("?", 0)
        text = src.get(f, [""] * (n + 1))[n - 1].strip()[:90] if n else ""
        stalls = ", ".join(f"{k[6:]}:{v}" for k, v in sorted(e[3].items(), key=lambda kv: -kv[1])[:3])
        print(f"  {100 * e[0] / max(tot_i, 1):5.1f}  {100 * e[1] / max(tot_s, 1):5.1f}  {f}:{n:<4d} {stalls} | {text}")


if __name__ == "__main__":
    main()
