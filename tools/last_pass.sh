#!/bin/bash
# last pass of the round: the GPU suite on the final host program, then the default bench line with the corrected traffic.json
o=gpurun_out; mkdir -p $o
timeout 600 python -m pytest tests -m gpu -x -q > $o/y_pytest.log 2>&1; echo "pytest rc=$?" >> $o/y_pytest.log; tail -3 $o/y_pytest.log
t0=$(date +%s); timeout 600 python bench.py > $o/y_bench.json 2> $o/y_bench.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"
python - <<P
import json
d=json.loads(open("$o/y_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "packed", d["e2e_packed"]["value"], "ragged", d.get("e2e_ragged"))
r=d["roofline"]; print("roofline", {k:r.get(k) for k in ("frac","dram_frac","own_floor_bytes","own_floor_frac","traffic")})
for c in d.get("configs",[]): print(c.get("config"), c.get("value"), c.get("ms_per_step"), c.get("fields_differing_vs_reference"), {k:(c.get("roofline") or {}).get(k) for k in ("frac","dram_frac")}, c.get("error"))
v=d.get("verify",{}); print("verify", v.get("value"), {k:(v.get("roofline") or {}).get(k) for k in ("frac","dram_frac")})
c=d.get("cli",{}); print("cli", {k:c.get(k) for k in ("ours_s","ours_runs_s","reference_s","outputs_identical","speedup")}); [print("  ",l) for l in c.get("ours_stages") or []]
P
