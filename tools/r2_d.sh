#!/bin/bash
# round-2 pass D: parity, then kernel-only timings of the parked/take-over pipeline and its knobs
o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/d_pytest.log 2>&1; echo "pytest rc=$?" >> $o/d_pytest.log; tail -4 $o/d_pytest.log
run() { # name workload env...
  local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/d_$name.json 2> $o/d_$name.err; echo "$name rc=$? $(cat $o/d_$name.json | cut -c1-160)"
}
run stress pe_stress WALT_X=0
run stress_tb3 pe_stress WALT_TAKE_BLOCKS=3
run stress_nopw pe_stress WALT_PAIR_WIDE=0
run stress_nohs pe_stress WALT_HEAP_SMEM=0
run stress_nolit pe_stress WALT_LIT=0
run se se WALT_X=0
run se_nolit se WALT_LIT=0
run se_tb3 se WALT_TAKE_BLOCKS=3
run se_d0 se WALT_DEFER=0
run pe pe WALT_X=0
run pe_d0 pe WALT_DEFER=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|lit_kernel" -c 40 --csv --log-file $o/d_se_launches.csv \
  python bench.py --workload se --steps 2 --warmup 3 --no-cpu --no-e2e > $o/d_se_l.log 2>&1; echo "se launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel|lit_kernel" -c 64 --csv --log-file $o/d_stress_launches.csv \
  python bench.py --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e > $o/d_stress_l.log 2>&1; echo "stress launches rc=$?"
python - <<P
import csv
for f in ("d_se_launches.csv","d_stress_launches.csv"):
    rows=[r for r in csv.reader(open("$o/"+f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
    for r in rows[1:17]:
        print(f[:8], r[ki][:60], r[gi], r[vi])
P
