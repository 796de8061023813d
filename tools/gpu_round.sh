#!/bin/bash
# tools/gpu_round.sh TAG : the standard GPU pass of a round, most important artefacts first --
#   parity tests, smoke; DRAM traffic of every mapping kernel of one device-resident step per workload ->
#   profiles/traffic.json (tools/make_traffic.py, on the box, so that the bench line that follows carries it);
#   the default bench line (all five configurations, verification micro-benchmark, cli leg); the reference arm;
#   full ncu captures of the kernels that matter, reduced to text here (the .ncu-rep files do not travel back:
#   gpurun_out/ is limited to 64 MiB); compute-sanitizer logs.
t=${1:-r}; o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q > $o/${t}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${t}_pytest.log; tail -3 $o/${t}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $o/${t}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $o/${t}_smoke.log

# DRAM bytes + duration of every mapping kernel (3 warm-up steps + the timed one + one whose counters are read)
K='regex:se_map|pe_|pair_kernel|verify_kernel|fold_kernel'
for wl in se pe_stress verify pe se_ag; do
  t0=$(date +%s)
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "$K" --csv \
    --log-file $o/${t}_traffic_$wl.csv python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu --no-e2e --no-configs > $o/${t}_traffic_$wl.log 2>&1
  echo "traffic $wl rc=$? in $(( $(date +%s) - t0 )) s"
done
python tools/make_traffic.py $o/$t | tee $o/${t}_traffic_summary.txt
cp profiles/traffic.json $o/${t}_traffic.json

t0=$(date +%s); timeout 1200 python bench.py > $o/${t}_bench.json 2> $o/${t}_bench.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"
python - <<P
import json
try:
    d=json.loads(open("$o/${t}_bench.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "packed", d["e2e_packed"]["value"])
    print("parity", d["parity_check"]); r=d["roofline"]; print("roofline", {k:r.get(k) for k in ("frac","dram_frac","own_floor_bytes","own_floor_frac","traffic")})
    for c in d.get("configs",[]): print(c.get("config"), c.get("value"), c.get("ms_per_step"), c.get("fields_differing_vs_reference"), {k:(c.get("roofline") or {}).get(k) for k in ("frac","dram_frac")}, c.get("error"), c.get("wall_s"))
    v=d.get("verify",{}); print("verify", v.get("value"), {k:(v.get("roofline") or {}).get(k) for k in ("frac","dram_frac")}, v.get("parity_check"), v.get("variants"), v.get("error"))
    c=d.get("cli",{}); print("cli", {k:c.get(k) for k in ("ours_s","reference_s","outputs_identical","speedup")}); [print("  ",l) for l in c.get("ours_stages") or []]
except Exception as ex:
    print("bench summary failed:", ex)
P

# full captures -> text: the single-end step (park / verify / fold / take), verify_kernel on its micro-benchmark, one
# paired-end stress step
full() {  # name, kernel filter, skip, count, bench arguments...
  local name=$1 k=$2 s=$3 c=$4; shift 4
  local t0=$(date +%s)
  timeout 500 ncu --set full --clock-control none --import-source on -k "$k" -s $s -c $c -o $o/${t}_$name -f python bench.py "$@" > $o/${t}_ncu_$name.log 2>&1
  echo "ncu $name rc=$? in $(( $(date +%s) - t0 )) s"
  if [ -f $o/${t}_$name.ncu-rep ]; then
    python tools/ncu_summary.py full $o/${t}_$name.ncu-rep > $o/${t}_${name}_counters.txt 2>&1
    ncu -i $o/${t}_$name.ncu-rep --page details > $o/${t}_${name}_details.txt 2>&1
    python tools/ncu_source_hot.py $o/${t}_$name.ncu-rep 40 > $o/${t}_${name}_source_hot.txt 2>&1
    rm -f $o/${t}_$name.ncu-rep
  fi
}
full se_step "$K" 12 4 --workload se --steps 1 --warmup 3 --no-cpu --no-e2e --no-configs
full verify_kernel regex:verify_kernel 3 1 --workload verify --steps 1 --warmup 3 --no-cpu
full stress_step "$K" 30 10 --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e --no-configs

t0=$(date +%s); timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${t}_reference.json 2> $o/${t}_reference.err; echo "reference rc=$? in $(( $(date +%s) - t0 )) s"; head -c 400 $o/${t}_reference.json; echo

# compute-sanitizer over the parity tests of both mapping paths, the repeat path (park / verify / fold / take), the groups
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "se_golden or pe_golden or repeats or edge_golden or group_and_clone" > $o/${t}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $o/${t}_memcheck.log | tail -3
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "se_golden or pe_golden or repeats" > $o/${t}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $o/${t}_racecheck.log | tail -3
rm -f $o/*.ncu-rep
du -sh $o
