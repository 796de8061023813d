#!/bin/bash
t=t; o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/${t}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${t}_pytest.log; grep -E "^(FAILED|ERROR)|passed|failed" $o/${t}_pytest.log | tail -5; grep -E "^E  " $o/${t}_pytest.log | head -12
timeout 900 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -x -q \
  -k "se_golden or pe_golden or repeats" > $o/${t}_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $o/${t}_racecheck.log | tail -3
grep -E "hazard|Race reported|at .*\+0x|walt_" $o/${t}_racecheck.log | head -40
