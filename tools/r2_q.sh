#!/bin/bash
# two-GPU pass: tests that need two devices, PCIe ceiling at 1 and 2 GPUs, the group API, the bench under torchrun
o=gpurun_out; mkdir -p $o
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -m gpu -x -q -k "two_physical or group_and_clone or sharded or clamped" > $o/q_pytest.log 2>&1; echo "pytest rc=$?" >> $o/q_pytest.log; tail -3 $o/q_pytest.log
for n in 1 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_probe.py pcie 2>/dev/null | grep probe
done
timeout 600 python tools/multi_gpu_probe.py group 2 2> $o/q_group.err | grep probe; tail -2 $o/q_group.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $o/q_bench2.json 2> $o/q_bench2.err; echo "bench2 rc=$?"
python - <<P
import json
d=json.loads(open("$o/q_bench2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "packed", d["e2e_packed"]["value"], d["parity_check"])
P
WALT_CLI_GPUS=2 timeout 600 python bench.py --workload cli --makedb-genome-mb 0 > $o/q_cli2.json 2> $o/q_cli2.err; echo "cli2 rc=$?"; python -c "
import json;d=json.loads(open('$o/q_cli2.json').read().strip().splitlines()[-1])['cli'];print({k:d.get(k) for k in ('gpus','ours_s','reference_s','outputs_identical','speedup','ours_stages')})"
