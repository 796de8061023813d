#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "repeats or pe_golden or se_golden or edge" > $o/m_pytest.log 2>&1; echo "pytest rc=$?" >> $o/m_pytest.log; tail -3 $o/m_pytest.log
run() { local name=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu --no-e2e --steps 3 > $o/m_$name.json 2> $o/m_$name.err; echo "$name rc=$? $(cat $o/m_$name.json | cut -c1-420)"
}
run stress pe_stress WALT_X=0
timeout 600 python bench.py --workload verify --steps 5 > $o/m_verify.json 2> $o/m_verify.err; echo "verify rc=$?"; cat $o/m_verify.json | cut -c400-1400; tail -2 $o/m_verify.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"verify_kernel" -s 4 -c 1 -o $o/m_verify_ncu -f \
  python bench.py --workload verify --steps 2 --warmup 3 --no-cpu > $o/m_verify_ncu.log 2>&1; echo "verify ncu rc=$?"
