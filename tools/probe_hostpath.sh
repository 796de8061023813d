#!/bin/bash
# host text path on the GPU box: file system of the work directory, write paths, the cli leg with part-size variants, the GPU suite
o=gpurun_out; mkdir -p $o
{ df -T /tmp /dev/shm . ; nproc; lscpu | grep -E "Model name|Flags" | cut -c1-300; free -g | head -2; } > $o/v_box.txt 2>&1
g++ -O2 -pthread -o /tmp/write_paths tools/probe/write_paths.cpp
for d in /tmp /dev/shm; do for m in "0 1" "1 8" "1 24" "2 8" "2 24"; do /tmp/write_paths $d/wp.bin $m 3000; done; done >> $o/v_box.txt 2>&1
cat $o/v_box.txt
WALT_CLI_VARIANTS="WALT_PART_READS=100000000;WALT_PART_READS=262144;WALT_PART_READS=2097152;WALT_BENCH_NOP=1" timeout 900 python bench.py --workload cli > $o/v_cli.json 2> $o/v_cli.err; echo "cli rc=$?"
python - <<P
import json
d=json.loads(open("$o/v_cli.json").read().strip().splitlines()[-1]); c=d["cli"]
print({k:c.get(k) for k in ("ours_s","ours_runs_s","reference_s","outputs_identical","speedup","setup_s")})
for l in c.get("ours_stages") or []: print(l)
for v in c.get("variants",[]): print(v["env"], v["s"], v["identical"]); [print("   ",l) for l in v["stages"]]
P
for k in 1 2; do timeout 900 python -m pytest tests -m gpu -x -q > $o/v_pytest$k.log 2>&1; echo "pytest$k rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" $o/v_pytest$k.log | tail -3; done
grep -h -E "^E  " $o/v_pytest*.log | head -10
