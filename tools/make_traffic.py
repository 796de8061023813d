#!/usr/bin/env python
"""profiles/traffic.json from the per-launch DRAM counters of one device-resident step per workload.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k 'regex:se_map|pe_|pair_kernel|verify_kernel|fold_kernel' --csv --log-file TAG_traffic_<kind>.csv \
        python bench.py --workload <kind> --steps 1 --warmup 3 --no-cpu --no-e2e

(tools/gpu_round.sh) runs 3 warm-up steps, the timed step and one more whose work counters are read.  A
single-end step starts with the parking kernel se_map_kernel<.., 1>, a paired-end step ends with pair_kernel;
the step that is summed is the FOURTH one (the timed one; all steps do the same work; a paired-end step
is several sub-launches of pairs, each with its own pair_kernel).  For `verify` only
the verify_kernel launches of that step count (its roofline line divides by their device time alone).

    python tools/make_traffic.py gpurun_out/TAG          -> writes profiles/traffic.json, prints a summary

Every entry carries the hash of the kernel sources it was measured on (bench.kernel_source_hash); bench.py
reports roofline.traffic / dram_frac only while that hash still matches."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KINDS = ("se", "se_ag", "pe", "pe_stress", "verify")
N_STEPS = 5    # --warmup 3 --steps 1 and the step whose work counters are read


def launches(path):
    """-> [{"name", "ns", "read", "write"}] in launch order"""
    rows, hdr, out = list(csv.reader(open(path, errors="replace"))), None, {}
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        e = out.setdefault(int(d["ID"]), {"name": d["Kernel Name"], "ns": 0.0, "read": 0.0, "write": 0.0})
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"].lower()
        m = d["Metric Name"]
        if m == "gpu__time_duration.sum":
            e["ns"] = v * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(u, 1.0)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
            e["read" if "read" in m else "write"] = v * scale
    return [out[k] for k in sorted(out)]


def steps_of(ls, paired):
    """cut the launch list into steps"""
    steps, cur = [], []
    for l in ls:
        n = l["name"]
        if not paired and re.search(r"se_map_kernel<\d+, \d+, 1>", n) and cur:
            steps.append(cur)
            cur = []
        cur.append(l)
        if paired and n.startswith("pair_kernel"):
            steps.append(cur)
            cur = []
    if cur:
        steps.append(cur)
    return steps


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "")


def main(prefix):
    sha = bench.kernel_source_hash()
    out = {"_doc": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum, ncu --clock-control none) of all mapping kernels of one "
                   "device-resident step of `bench.py --workload <kind>` at full size (for `verify`: of its verify_kernel launches), "
                   "made by tools/make_traffic.py from the per-launch lists named in `source`; valid for the kernel sources with "
                   "this source_sha16 only (bench.kernel_source_hash)"}
    for kind in KINDS:
        path = f"{prefix}_traffic_{kind}.csv"
        if not os.path.exists(path):
            print(f"{kind}: no {path}")
            continue
        ls = launches(path)
        paired = kind in ("pe", "pe_stress")
        steps = steps_of(ls, paired)
        if paired:
            # a paired-end device call runs as several sub-launches of pairs (each ends with its pair_kernel):
            # the capture holds N_STEPS whole steps, so every step is len / N_STEPS consecutive sub-launches
            if len(steps) % N_STEPS:
                print(f"{kind}: {len(steps)} sub-launches do not make {N_STEPS} steps in {path}")
                continue
            per_step = len(steps) // N_STEPS
            steps = [sum(steps[i * per_step:(i + 1) * per_step], []) for i in range(N_STEPS)]
        if len(steps) < 4:
            print(f"{kind}: only {len(steps)} steps in {path}")
            continue
        step = steps[3]
        if kind == "verify":
            step = [l for l in step if "verify_kernel" in l["name"]]
        gmb, n, _ = bench.FULL_SIZE[kind]
        per = {}
        for l in step:
            k = per.setdefault(short(l["name"]), {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
            k["launches"] += 1
            k["ms"] += l["ns"] / 1e6
            k["dram_bytes"] += l["read"] + l["write"]
        for k in per.values():
            k["ms"] = round(k["ms"], 4)
            k["dram_bytes"] = int(k["dram_bytes"])
        total = int(sum(l["read"] + l["write"] for l in step))
        out[kind] = {"reads_per_step": n, "genome_mb": gmb, "source_sha16": sha, "dram_bytes_per_step": total,
                     "dram_bytes_read": int(sum(l["read"] for l in step)), "dram_bytes_written": int(sum(l["write"] for l in step)),
                     "kernel_ms_under_ncu": round(sum(l["ns"] for l in step) / 1e6, 4), "launches_per_step": len(step),
                     "steps_in_capture": len(steps), "kernels": per,
                     "source": f"profiles/{os.path.basename(prefix)}_traffic_{kind}.csv"}
        print(f"{kind}: {len(steps)} steps, step 4: {len(step)} launches, {total / 1e9:.3f} GB, "
              f"{out[kind]['kernel_ms_under_ncu']:.3f} ms serialised, {total / n:.0f} B per read")
    if len(out) > 1:
        json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
        print("wrote profiles/traffic.json for sources", sha)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
