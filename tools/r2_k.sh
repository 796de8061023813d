#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "repeats or pe_golden or se_golden" > $o/k_pytest.log 2>&1; echo "pytest rc=$?" >> $o/k_pytest.log; tail -3 $o/k_pytest.log
run() { local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/k_$name.json 2> $o/k_$name.err; echo "$name rc=$? $(cat $o/k_$name.json | cut -c1-330)"
}
run stress pe_stress WALT_X=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|verify_kernel|fold_kernel" -c 44 --csv --log-file $o/k_stress_launches.csv \
  python bench.py --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e > $o/k_stress_l.log 2>&1; echo "stress launches rc=$?"
python - <<P
import csv
for f in ("k_stress_launches.csv",):
    rows=[r for r in csv.reader(open("$o/"+f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
    for r in rows[1:12]:
        print(f[:8], r[ki][:60], r[gi], r[vi])
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"verify_kernel|fold_kernel" -s 12 -c 2 -o $o/k_vf -f \
  python bench.py --workload pe_stress --genome-mb 1000 --reads 1000000 --steps 1 --warmup 3 --no-cpu --no-e2e > $o/k_vf_ncu.log 2>&1; echo "vf ncu rc=$?"
