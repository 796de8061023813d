#!/bin/bash
# tools/gpu_round.sh TAG : the standard GPU pass of a round -- parity tests, smoke, the default bench line (all five
# configurations, verification micro-benchmark, cli leg), the reference arm, DRAM traffic of every mapping kernel of one
# device-resident step per workload (what profiles/traffic.json is made from), full ncu captures of the kernels that
# matter, compute-sanitizer logs.
t=${1:-r}; o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/${t}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${t}_pytest.log; tail -3 $o/${t}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $o/${t}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $o/${t}_smoke.log
t0=$(date +%s); timeout 1500 python bench.py > $o/${t}_bench.json 2> $o/${t}_bench.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"
t0=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${t}_reference.json 2> $o/${t}_reference.err; echo "reference rc=$? in $(( $(date +%s) - t0 )) s"
# DRAM bytes + duration of every mapping kernel of the last device-resident step (3 warm-up + 1 timed + 1 counted step)
K='regex:se_map|pe_|pair_kernel|verify_kernel|fold_kernel'
for wl in se se_ag pe pe_stress verify; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" --csv \
    --log-file $o/${t}_traffic_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu --no-e2e > $o/${t}_traffic_$wl.log 2>&1; echo "traffic $wl rc=$?"
done
# full captures: the single-end step (park / verify / fold / take), the verification kernel on its micro-benchmark,
# one paired-end stress chunk
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 16 -c 4 -o $o/${t}_se_full -f \
  python bench.py --workload se --steps 1 --warmup 3 --no-cpu --no-e2e > $o/${t}_ncu_se.log 2>&1; echo "ncu se rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:verify_kernel -s 4 -c 1 -o $o/${t}_verify_full -f \
  python bench.py --workload verify --steps 2 --warmup 3 --no-cpu > $o/${t}_ncu_verify.log 2>&1; echo "ncu verify rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 30 -c 10 -o $o/${t}_stress_full -f \
  python bench.py --workload pe_stress --genome-mb 1000 --reads 1000000 --steps 1 --warmup 3 --no-cpu --no-e2e > $o/${t}_ncu_stress.log 2>&1; echo "ncu stress rc=$?"
# compute-sanitizer: memcheck and racecheck over the parity tests of both mapping paths, the repeat path and the builder
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "se_golden or pe_golden or repeats or edge_golden or group_and_clone or makedb_matches_oracle" > $o/${t}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $o/${t}_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "se_golden or pe_golden or repeats" > $o/${t}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $o/${t}_racecheck.log | tail -3
python - <<P
import json
d=json.loads(open("$o/${t}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "packed", d["e2e_packed"]["value"])
print("parity", d["parity_check"]); r=d["roofline"]; print("roofline", {k:r[k] for k in ("frac","dram_frac","own_floor_bytes","own_floor_frac","traffic")})
for c in d.get("configs",[]): print(c.get("config"), c.get("value"), c.get("ms_per_step"), c.get("fields_differing_vs_reference"), (c.get("roofline") or {}).get("frac"), c.get("error"), c.get("wall_s"))
v=d.get("verify",{}); print("verify", v.get("value"), (v.get("roofline") or {}).get("frac"), v.get("parity_check"), v.get("variants"), v.get("error"))
print("cli", {k:d.get("cli",{}).get(k) for k in ("ours_s","reference_s","outputs_identical","speedup")})
print("ref", open("$o/${t}_reference.json").read()[:300])
P
