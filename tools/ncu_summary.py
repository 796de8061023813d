#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: `launches <csv>` -> per-kernel time shares;
`full <ncu-rep>` -> the counters DESIGN.md / bench.py quote (dram bytes, hit rates, stalls)."""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "smsp__cycles_active.avg"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"])[:90]
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        ms = v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v if u.startswith("ms") else v * 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.2f} ms total (cold-cache, serialised)")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:12.3f} ms {n:5d}x {100 * ms / tot:6.2f}%  avg {ms / n:10.4f} ms  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = dict(zip(h, v))
        print(f"## {d.get('Kernel Name', '?')[:100]}")
        for i, n in enumerate(h):
            if n in KEEP or ("issue_stalled" in n and n.endswith("per_issue_active.ratio")):
                print(f"{n:80s} {v[i]:>20s} {u[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
