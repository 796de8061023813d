#!/bin/bash
# full ncu captures: park, take-over, heap and pair kernels of the stress workload (reduced size, one chunk per step)
o=gpurun_out; mkdir -p $o
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pe_|pair_kernel" -s 18 -c 6 -o $o/c_stress_take -f \
  python bench.py --workload pe_stress --genome-mb 1000 --reads 1000000 --steps 1 --warmup 3 --no-cpu --no-e2e > $o/c_stress.log 2>&1; echo "stress rc=$?"
ls -la $o/c_*.ncu-rep
