#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/e_pytest.log 2>&1; echo "pytest rc=$?" >> $o/e_pytest.log; tail -4 $o/e_pytest.log
run() { local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/e_$name.json 2> $o/e_$name.err; echo "$name rc=$? $(cat $o/e_$name.json | cut -c1-330)"
}
run stress pe_stress WALT_X=0
run stress_tb3 pe_stress WALT_TAKE_BLOCKS=3
run se se WALT_X=0
run se_tb3 se WALT_TAKE_BLOCKS=3
run pe pe WALT_X=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|lit_kernel" -c 30 --csv --log-file $o/e_se_launches.csv \
  python bench.py --workload se --steps 2 --warmup 3 --no-cpu --no-e2e > $o/e_se_l.log 2>&1; echo "se launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|lit_kernel" -c 32 --csv --log-file $o/e_stress_launches.csv \
  python bench.py --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e > $o/e_stress_l.log 2>&1; echo "stress launches rc=$?"
python - <<P
import csv
for f in ("e_se_launches.csv","e_stress_launches.csv"):
    rows=[r for r in csv.reader(open("$o/"+f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
    for r in rows[1:9]:
        print(f[:8], r[ki][:60], r[gi], r[vi])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pe_log_kernel" -s 25 -c 2 -o $o/e_stress_take -f \
  python bench.py --workload pe_stress --genome-mb 1000 --reads 1000000 --steps 1 --warmup 3 --no-cpu --no-e2e > $o/e_stress_ncu.log 2>&1; echo "stress ncu rc=$?"
