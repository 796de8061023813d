#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/o_pytest.log 2>&1; echo "pytest rc=$?" >> $o/o_pytest.log; tail -4 $o/o_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $o/o_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $o/o_smoke.log
t0=$(date +%s)
timeout 1500 python bench.py > $o/o_bench.json 2> $o/o_bench.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -3 $o/o_bench.err
python - <<P
import json
d=json.loads(open("$o/o_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "packed", d["e2e_packed"]["value"])
print("parity", d["parity_check"])
r=d["roofline"]; print("roofline", {k:r[k] for k in ("frac","dram_frac","own_floor_bytes","own_floor_frac","traffic")})
print("cpu", d["cpu_baseline"])
for c in d.get("configs",[]): print(c.get("config"), c.get("value"), c.get("ms_per_step"), c.get("fields_differing_vs_reference"), (c.get("roofline") or {}).get("frac"), c.get("error"), c.get("wall_s"), (c.get("cli") or {}).get("outputs_identical"), (c.get("reference") or {}).get("value"))
v=d.get("verify",{}); print("verify", v.get("value"), (v.get("roofline") or {}).get("frac"), v.get("parity_check"), v.get("error"))
print("cli", {k:d.get("cli",{}).get(k) for k in ("ours_s","reference_s","outputs_identical","speedup","ours_stages")})
P
