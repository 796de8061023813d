#!/bin/bash
# where the start-up time of the walt program goes: cli leg without and with a sync after the set-up, repeated runs
o=gpurun_out; mkdir -p $o
for sy in 0 1; do
WALT_CLI_SYNC=$sy WALT_CLI_VARIANTS="R=1;R=2;CUDA_MODULE_LOADING=EAGER" timeout 210 python bench.py --workload cli --makedb-genome-mb 0 > $o/x_cli_sync$sy.json 2> $o/x_cli_sync$sy.err; echo "cli sync=$sy rc=$?"
python - <<P
import json
d=json.loads(open("$o/x_cli_sync$sy.json").read().strip().splitlines()[-1]); c=d["cli"]
print({k:c.get(k) for k in ("ours_s","ours_runs_s","reference_s","outputs_identical","speedup","setup_s")})
for l in c.get("ours_stages") or []: print(l)
for v in c.get("variants",[]): print(v["env"], v["s"], v["identical"]); [print("   ",l) for l in v["stages"]]
P
done
