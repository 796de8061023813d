#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/j_pytest.log 2>&1; echo "pytest rc=$?" >> $o/j_pytest.log; tail -4 $o/j_pytest.log
if grep -q failed $o/j_pytest.log; then
  timeout 900 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_parity.py -x -q -k "repeats" > $o/j_san.log 2>&1; echo "san rc=$?"
  grep -E "Invalid|at 0x|Address|walt_core|walt_engine|ERROR SUMMARY|passed|failed|Error" $o/j_san.log | head -30
fi
run() { local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/j_$name.json 2> $o/j_$name.err; echo "$name rc=$? $(cat $o/j_$name.json | cut -c1-330)"
}
run stress pe_stress WALT_X=0
run stress_noflat pe_stress WALT_FLAT=0
run se se WALT_X=0
run pe pe WALT_X=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|verify_kernel|fold_kernel" -c 44 --csv --log-file $o/j_stress_launches.csv \
  python bench.py --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e > $o/j_stress_l.log 2>&1; echo "stress launches rc=$?"
python - <<P
import csv
for f in ("j_stress_launches.csv",):
    rows=[r for r in csv.reader(open("$o/"+f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
    for r in rows[1:12]:
        print(f[:8], r[ki][:60], r[gi], r[vi])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"verify_kernel" -s 6 -c 1 -o $o/j_verify -f \
  python bench.py --workload pe_stress --genome-mb 1000 --reads 1000000 --steps 1 --warmup 3 --no-cpu --no-e2e > $o/j_verify_ncu.log 2>&1; echo "verify ncu rc=$?"
