#!/bin/bash
# tools/gpu_round.sh TAG : the standard GPU pass -- parity tests, the bench lines of every config
# with their side legs, the reference arm, the ncu launch list of the bench command and one full
# capture of each mapping kernel.
t=$1; o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q > $o/${t}_pytest.log 2>&1; echo "pytest rc=$?" >> $o/${t}_pytest.log
tail -3 $o/${t}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $o/${t}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $o/${t}_smoke.log
timeout 900 python bench.py > $o/${t}_se.json 2> $o/${t}_se.err; echo "se rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${t}_reference.json 2> $o/${t}_reference.err; echo "reference rc=$?"
timeout 600 python bench.py --workload se_ag > $o/${t}_se_ag.json 2> $o/${t}_se_ag.err; echo "se_ag rc=$?"
timeout 600 python bench.py --workload pe > $o/${t}_pe.json 2> $o/${t}_pe.err; echo "pe rc=$?"
timeout 900 python bench.py --workload pe_stress > $o/${t}_pe_stress.json 2> $o/${t}_pe_stress.err; echo "pe_stress rc=$?"
# launch list of the bench command, restricted to the mapping kernels (the index build in front of them is set-up)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"se_map|pe_|pair_kernel" -c 400 --csv --log-file $o/${t}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > $o/${t}_ncu_launches.log 2>&1; echo "launches rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file $o/${t}_launches_setup.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1; echo "setup launches rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:se_map -s 3 -c 1 -o $o/${t}_se_full -f \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $o/${t}_ncu_se.log 2>&1; echo "ncu se rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pe_|pair_kernel" -s 12 -c 4 -o $o/${t}_pe_full -f \
  python bench.py --workload pe --genome-mb 1000 --steps 2 --warmup 3 --no-cpu --no-e2e > $o/${t}_ncu_pe.log 2>&1; echo "ncu pe rc=$?"
python - <<P
import json
for w in ("se","reference","se_ag","pe","pe_stress"):
    try:
        d=json.loads(open("$o/${t}_%s.json"%w).read().strip().splitlines()[-1])
        print(w, d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("e2e_packed") or {}).get("value"), d.get("roofline") and d["roofline"]["frac"], d.get("parity_check"), (d.get("cli") or {}).get("speedup"))
    except Exception as ex: print(w, "ERR", ex)
P
