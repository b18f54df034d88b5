#!/bin/bash
mkdir -p gpurun_out
export KID_BENCH_TRACE=1
timeout -k 5 120 python -m pytest tests/test_multirank_gpu.py -m gpu -q -k "uneven or nccl or overlapped" > gpurun_out/r2s_test.log 2>&1; tail -3 gpurun_out/r2s_test.log | cut -c1-300
run() { tag=$1; shift; timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 20 --warmup 5 "$@" > gpurun_out/r2s_$tag.json 2> gpurun_out/r2s_$tag.err; echo "$tag rc=$?"; grep "bench rank" gpurun_out/r2s_$tag.err | tail -2 | cut -c1-250; }
run c
python - <<PY
import json
d=json.loads(open("gpurun_out/r2s_c.json").read().strip().splitlines()[-1])
print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "kern", d["roofline"]["kernel_ms"], "parity ok", d.get("parity_nccl",{}).get("ok"), "e2e", d.get("e2e",{}).get("ms_per_step"), d.get("e2e",{}).get("unpipelined_ms_per_step"), d["config"]["per_rank"], d["config"]["migration_ms_per_step"], d["config"]["sort_ms_per_call"])
PY
