#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2q_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -3 gpurun_out/r2q_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2q_bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "kern", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "unpipelined", d["e2e"]["unpipelined_ms_per_step"], "weak", d["weak_base"]["ms_per_step"], d["weak_base"]["kernel_ms"])
PY
