#!/usr/bin/env python
"""Warp-instructions and stall samples of an ncu report (--import-source on) aggregated by
source function: `ncu_source_regions.py rep.ncu-rep file.cuh [file.cu ...]` maps each line to the
nearest preceding function-like definition in the given sources."""
import csv
import re
import subprocess
import sys

rep, srcs = sys.argv[1], sys.argv[2:]
funcs = {}
for path in srcs:
    name = path.split("/")[-1]
    cur = "?"
    table = []
    for i, line in enumerate(open(path), 1):
        m = re.match(r"^(?:template.*>\s*)?(?:WALT_HD|static|__global__|__device__|inline)?[\w\s\*&:<>,]*?\b(\w+)\s*\([^;]*$", line)
        if m and not line.startswith(" ") and not line.startswith("//") and "(" in line and not line.startswith("#"):
            cur = m.group(1)
        table.append(cur)
    funcs[name] = table
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, agg = "?", None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 5 or not r[0]:
        continue
    d = dict(zip(hdr, r))
    try:
        inst = int(d["Instructions Executed"]); samp = int(d["# Samples"]); thr = int(d["Thread Instructions Executed"])
        ln = int(r[0])
    except (ValueError, KeyError):
        continue
    t = funcs.get(cur_file)
    fn = t[ln - 1] if t and ln - 1 < len(t) else cur_file
    a = agg.setdefault(fn, [0, 0, 0])
    a[0] += inst; a[1] += samp; a[2] += thr
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp-instructions {ti:,}  stall samples {ts:,}")
for fn, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] * 1000 < ti and a[1] * 1000 < ts:
        continue
    print(f"{100*a[0]/ti:5.1f}% inst {100*a[1]/max(ts,1):5.1f}% stall  thr/inst {a[2]/max(a[0],1):5.1f}  {fn}")
