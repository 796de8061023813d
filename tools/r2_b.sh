#!/bin/bash
# launch lists (ncu, serialised) of the SE and PE-stress device steps: which kernel takes what
o=gpurun_out; mkdir -p $o
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel" -c 40 --csv --log-file $o/b_se_launches.csv \
  python bench.py --workload se --steps 2 --warmup 3 --no-cpu --no-e2e > $o/b_se.log 2>&1; echo "se rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel" -c 60 --csv --log-file $o/b_stress_launches.csv \
  python bench.py --workload pe_stress --steps 1 --warmup 3 --no-cpu --no-e2e > $o/b_stress.log 2>&1; echo "stress rc=$?"
python - <<P
import csv
for f in ("b_se_launches.csv","b_stress_launches.csv"):
    rows=[r for r in csv.reader(open("$o/"+f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
    for r in rows[1:41]:
        print(f[:8], r[ki][:70], r[gi], r[vi])
P
