#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/h_pytest.log 2>&1; echo "pytest rc=$?" >> $o/h_pytest.log; tail -4 $o/h_pytest.log
if grep -q failed $o/h_pytest.log; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -x -q -k "se_packed or se_empty or pe_golden" > $o/h_san.log 2>&1; echo "san rc=$?"
  grep -E "Invalid|at 0x|Address|walt_core|walt_engine|ERROR SUMMARY|passed|failed|Error" $o/h_san.log | head -30
fi
run() { local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/h_$name.json 2> $o/h_$name.err; echo "$name rc=$? $(cat $o/h_$name.json | cut -c1-330)"
}
run stress pe_stress WALT_X=0
run se se WALT_X=0
run se_d0 se WALT_DEFER=0
run pe pe WALT_X=0
run pe_d0 pe WALT_DEFER=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|lit_kernel" -c 32 --csv --log-file $o/h_se_launches.csv \
  python bench.py --workload se --steps 2 --warmup 3 --no-cpu --no-e2e > $o/h_se_l.log 2>&1; echo "se launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|lit_kernel" -c 40 --csv --log-file $o/h_stress_launches.csv \
  python bench.py --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e > $o/h_stress_l.log 2>&1; echo "stress launches rc=$?"
python - <<P
import csv
for f in ("h_se_launches.csv","h_stress_launches.csv"):
    rows=[r for r in csv.reader(open("$o/"+f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
    for r in rows[1:11]:
        print(f[:8], r[ki][:60], r[gi], r[vi])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pe_log_kernel" -s 33 -c 1 -o $o/h_stress_take -f \
  python bench.py --workload pe_stress --genome-mb 1000 --reads 1000000 --steps 1 --warmup 3 --no-cpu --no-e2e > $o/h_stress_ncu.log 2>&1; echo "stress ncu rc=$?"
