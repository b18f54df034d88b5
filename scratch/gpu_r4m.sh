#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r4m_n2.json 2> gpurun_out/r4m_n2.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r4m_n2.json").read().strip().splitlines()[-1])
print("N=2 ms/step", d["ms_per_step"], "value %.4g" % d["value"], "kern", d["roofline"]["kernel_ms"], "parity ok", d.get("parity_nccl",{}).get("ok"), "migrated", d.get("parity_nccl",{}).get("migrated"), "e2e", d.get("e2e",{}).get("ms_per_step"), d["config"]["migration_ms_per_step"], d["config"]["sort_ms_per_call"])
PY
