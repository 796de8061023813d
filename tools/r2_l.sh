#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/l_pytest.log 2>&1; echo "pytest rc=$?" >> $o/l_pytest.log; tail -3 $o/l_pytest.log
run() { local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/l_$name.json 2> $o/l_$name.err; echo "$name rc=$? $(cat $o/l_$name.json | cut -c1-420)"
}
run stress pe_stress WALT_X=0
run stress_noflat pe_stress WALT_FLAT=0
run se se WALT_X=0
run pe pe WALT_X=0
timeout 600 python bench.py --workload verify --steps 5 > $o/l_verify.json 2> $o/l_verify.err; echo "verify rc=$?"; cat $o/l_verify.json | cut -c1-1500; tail -2 $o/l_verify.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|verify_kernel|fold_kernel" -c 44 --csv --log-file $o/l_stress_launches.csv \
  python bench.py --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e > $o/l_stress_l.log 2>&1; echo "stress launches rc=$?"
python - <<P
import csv
for f in ("l_stress_launches.csv",):
    rows=[r for r in csv.reader(open("$o/"+f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
    for r in rows[1:12]:
        print(f[:8], r[ki][:60], r[gi], r[vi])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"verify_kernel" -s 4 -c 1 -o $o/l_verify_ncu -f \
  python bench.py --workload verify --steps 2 --warmup 3 --no-cpu > $o/l_verify_ncu.log 2>&1; echo "verify ncu rc=$?"
