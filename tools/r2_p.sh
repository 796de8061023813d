#!/bin/bash
o=gpurun_out; mkdir -p $o
for sl in 3 5 6; do
WALT_SE_SLOTS=$sl CHUNKS="524288,1048576" timeout 600 python tools/e2e_probe.py > $o/p_probe_s$sl.txt 2> $o/p_probe_s$sl.err; echo "slots $sl rc=$?"
grep -h e2e $o/p_probe_s$sl.txt | cut -c1-120
done
