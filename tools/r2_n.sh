#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "repeats or pe_golden or se_golden or edge" > $o/n_pytest.log 2>&1; echo "pytest rc=$?" >> $o/n_pytest.log; tail -3 $o/n_pytest.log
timeout 600 python bench.py --workload verify --steps 5 > $o/n_verify.json 2> $o/n_verify.err; echo "verify rc=$? $(cat $o/n_verify.json | cut -c400-1300)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"verify_kernel" -s 4 -c 1 -o $o/n_verify_ncu -f \
  python bench.py --workload verify --steps 2 --warmup 3 --no-cpu > $o/n_verify_ncu.log 2>&1; echo "verify ncu rc=$?"
