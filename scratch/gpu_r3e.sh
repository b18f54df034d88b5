#!/bin/bash
mkdir -p gpurun_out
export KID_BENCH_TRACE=1
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r3e_n4.json 2> gpurun_out/r3e_n4.err; echo "rc=$?"; grep "bench rank 0" gpurun_out/r3e_n4.err | tail -2 | cut -c1-200
python - <<PY
import json
d=json.loads(open("gpurun_out/r3e_n4.json").read().strip().splitlines()[-1])
print("N=4 ms/step", d["ms_per_step"], "value %.4g" % d["value"], "kern", d["roofline"]["kernel_ms"], "parity ok", d.get("parity_nccl",{}).get("ok"), "migrated", d.get("parity_nccl",{}).get("migrated"), "e2e", d.get("e2e",{}).get("ms_per_step"), d["config"]["per_rank"]["kernel_ms"], d["config"]["migration_ms_per_step"], d["config"]["sort_ms_per_call"])
PY
