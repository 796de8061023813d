mkdir -p gpurun_out
o=gpurun_out/s18.txt; : > $o
for envs in "X=1" "WALT_PE_SIDE=0"; do
  for w in pe pe_stress; do
    echo "## $envs $w" >> $o
    env $envs timeout 400 python bench.py --workload $w --no-cpu --no-e2e --steps 5 --warmup 3 >> $o 2>&1
  done
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pe_map|pair_kernel" -c 30 --csv --log-file gpurun_out/s18_pe_launches.csv python bench.py --workload pe --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/s18_ncu.log 2>&1
cat $o
