#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout -k 5 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e --no-parity > gpurun_out/r2v_$tag.json 2> gpurun_out/r2v_$tag.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2v_$tag.json").read().strip().splitlines()[-1])
    print("$tag ms/step %.4f kern %.4f mig %.4f sort/call %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["migration_ms_per_step"], d["config"]["sort_ms_per_call"]), d["config"]["per_rank"]["kernel_ms"])
except Exception as e: print("$tag FAILED", e)
PY
}
run tma
KID_NO_TMA=1 run notma
KID_NO_PIPELINE=1 run tma_nopipe
KID_SCATTER_DENSE=0 run tma_sparse
