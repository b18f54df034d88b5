#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2r_n2.json 2> gpurun_out/r2r_n2.err; tail -3 gpurun_out/r2r_n2.err | cut -c1-300; python - <<PY
import json
d=json.loads(open("gpurun_out/r2r_n2.json").read().strip().splitlines()[-1])
print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "kern", d["roofline"]["kernel_ms"], "parity", d.get("parity_nccl"), "e2e", d.get("e2e",{}).get("ms_per_step"), d["config"]["per_rank"])
PY
timeout 600 python -m pytest tests/test_multirank_gpu.py -m gpu -q -k nccl > gpurun_out/r2r_nccl_test.log 2>&1; tail -3 gpurun_out/r2r_nccl_test.log | cut -c1-300
