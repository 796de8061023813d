#!/bin/bash
# eight-GPU pass: the node's host<->device ceiling at 1/2/4/8 GPUs, one process driving 8 engines, the bench, walt -gpus 8
o=gpurun_out; mkdir -p $o
nvidia-smi -L | wc -l; nproc; free -g | head -2
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/multi_gpu_probe.py pcie 2>/dev/null | grep probe | tee -a $o/r_pcie.jsonl
done
for n in 4 8; do
timeout 900 python tools/multi_gpu_probe.py group $n 2> $o/r_group$n.err | grep probe | tee -a $o/r_group.jsonl; tail -1 $o/r_group$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > $o/r_bench8.json 2> $o/r_bench8.err; echo "bench8 rc=$?"
python - <<P
import json
d=json.loads(open("$o/r_bench8.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "packed", d["e2e_packed"]["value"], d["parity_check"])
P
WALT_CLI_GPUS=8 timeout 600 python bench.py --workload cli --makedb-genome-mb 0 > $o/r_cli8.json 2> $o/r_cli8.err; echo "cli8 rc=$?"; python -c "
import json;d=json.loads(open('$o/r_cli8.json').read().strip().splitlines()[-1])['cli'];print({k:d.get(k) for k in ('gpus','ours_s','reference_s','outputs_identical','speedup','ours_stages')})"
