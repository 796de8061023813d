#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/f_pytest.log 2>&1; echo "pytest rc=$?" >> $o/f_pytest.log; tail -4 $o/f_pytest.log
run() { local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/f_$name.json 2> $o/f_$name.err; echo "$name rc=$? $(cat $o/f_$name.json | cut -c1-330)"
}
run stress pe_stress WALT_X=0
run stress_l1 pe_stress WALT_LIT_LEVELS=1
run stress_l3 pe_stress WALT_LIT_LEVELS=3
run stress_noside pe_stress WALT_LIT_SIDE=0
run se se WALT_X=0
run se_l1 se WALT_LIT_LEVELS=1
run se_l3 se WALT_LIT_LEVELS=3
run se_d0 se WALT_DEFER=0
run pe pe WALT_X=0
run pe_d0 pe WALT_DEFER=0
run se_ag se_ag WALT_X=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|lit_kernel" -c 32 --csv --log-file $o/f_se_launches.csv \
  python bench.py --workload se --steps 2 --warmup 3 --no-cpu --no-e2e > $o/f_se_l.log 2>&1; echo "se launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|lit_kernel" -c 40 --csv --log-file $o/f_stress_launches.csv \
  python bench.py --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e > $o/f_stress_l.log 2>&1; echo "stress launches rc=$?"
python - <<P
import csv
for f in ("f_se_launches.csv","f_stress_launches.csv"):
    rows=[r for r in csv.reader(open("$o/"+f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
    for r in rows[1:11]:
        print(f[:8], r[ki][:60], r[gi], r[vi])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pe_log_kernel" -s 32 -c 3 -o $o/f_stress_take -f \
  python bench.py --workload pe_stress --genome-mb 1000 --reads 1000000 --steps 1 --warmup 3 --no-cpu --no-e2e > $o/f_stress_ncu.log 2>&1; echo "stress ncu rc=$?"
