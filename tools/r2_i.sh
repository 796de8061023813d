#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/i_pytest.log 2>&1; echo "pytest rc=$?" >> $o/i_pytest.log; tail -4 $o/i_pytest.log
# flow test of the default bench line at reduced size
WALT_BENCH_SCALE=0.05 timeout 900 python bench.py --genome-mb 155 --reads 500000 --cli-genome-mb 30 --cli-reads 200000 > $o/i_flow.json 2> $o/i_flow.err; echo "flow rc=$?"; tail -3 $o/i_flow.err
python - <<P
import json
d=json.loads(open("$o/i_flow.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["e2e_packed"]["value"], d["parity_check"], d["roofline"] and {k:d["roofline"][k] for k in ("frac","dram_frac","own_floor_bytes","own_floor_frac")})
for c in d.get("configs",[]): print(c.get("config"), c.get("value"), c.get("ms_per_step"), c.get("fields_differing_vs_reference"), (c.get("roofline") or {}).get("frac"), c.get("error"), c.get("wall_s"), (c.get("cli") or {}).get("outputs_identical"))
print(d.get("cli",{}).get("outputs_identical"), d.get("cli",{}).get("speedup"))
P
run() { local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/i_$name.json 2> $o/i_$name.err; echo "$name rc=$? $(cat $o/i_$name.json | cut -c1-330)"
}
run stress pe_stress WALT_X=0
run se se WALT_X=0
run pe pe WALT_X=0
run small se_small WALT_X=0
