#!/bin/bash
o=gpurun_out; mkdir -p $o
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "test_se_golden" > $o/san_se.log 2>&1; echo "rc=$?"
grep -E "Invalid|at 0x|by thread|Address|walt_core|walt_engine|ERROR SUMMARY" $o/san_se.log | head -60
