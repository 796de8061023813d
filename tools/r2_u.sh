#!/bin/bash
o=gpurun_out; mkdir -p $o
for k in 1 2 3; do
timeout 1200 python -m pytest tests -m gpu -q > $o/u_pytest$k.log 2>&1; echo "pytest$k rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $o/u_pytest$k.log | tail -4
done
grep -h -E "^E  |^FAILED" $o/u_pytest*.log | head -20
