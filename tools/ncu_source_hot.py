#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report captured with --import-source on:
instructions executed and stall samples aggregated by file:line (top N)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = "?"
hdr = None
agg = {}
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
    except (ValueError, KeyError):
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:90]])
    a[0] += inst; a[1] += samp; a[2] += thr
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot_i:,}  stall samples {tot_s:,}")
print("by instructions:")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot_i:5.1f}% inst {100*a[1]/max(tot_s,1):5.1f}% stall  thr/inst {a[2]/max(a[0],1):5.1f}  {f}:{l}  {a[3]}")
print("by stall samples:")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top // 2]:
    print(f"{100*a[0]/tot_i:5.1f}% inst {100*a[1]/max(tot_s,1):5.1f}% stall  thr/inst {a[2]/max(a[0],1):5.1f}  {f}:{l}  {a[3]}")
