#!/bin/bash
# round-2 pass A: parity of the parked/take-over path, then kernel-only timings of every workload with and without it
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q > $o/a_pytest.log 2>&1; echo "pytest rc=$?" >> $o/a_pytest.log; tail -4 $o/a_pytest.log
for d in 1 0; do
  WALT_DEFER=$d timeout 600 python bench.py --workload pe_stress --no-cpu --no-e2e --steps 3 > $o/a_stress_d$d.json 2> $o/a_stress_d$d.err; echo "stress d$d rc=$?"; cat $o/a_stress_d$d.json
done
for d in 1 0; do
  WALT_DEFER=$d timeout 600 python bench.py --workload se --no-cpu --no-e2e --steps 5 > $o/a_se_d$d.json 2> $o/a_se_d$d.err; echo "se d$d rc=$?"; cat $o/a_se_d$d.json
done
WALT_DEFER=1 timeout 600 python bench.py --workload pe --no-cpu --no-e2e --steps 5 > $o/a_pe_d1.json 2> $o/a_pe_d1.err; echo "pe rc=$?"; cat $o/a_pe_d1.json
