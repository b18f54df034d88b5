#!/bin/bash
mkdir -p gpurun_out
KID_BENCH_TRACE=1 timeout -k 5 500 python bench.py --bergs-per-gpu 100000000 --steps 20 --warmup 5 --no-cpu --no-e2e --no-weak-base > gpurun_out/r3a_1e8.json 2> gpurun_out/r3a_1e8.err; tail -4 gpurun_out/r3a_1e8.err | cut -c1-200
python - <<PY
import json
d=json.loads(open("gpurun_out/r3a_1e8.json").read().strip().splitlines()[-1])
print("1e8: ms/step", d["ms_per_step"], "value", d["value"], "kern", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "sort/call", d["config"]["sort_ms_per_call"], d["clocks"])
PY
