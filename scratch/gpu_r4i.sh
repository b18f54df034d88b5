#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_interactions_gpu.py tests/test_rk_interactions_gpu.py tests/test_footloose_gpu.py "tests/test_multirank_gpu.py::test_interacting_bergs_across_ranks" -m gpu -q --no-header > gpurun_out/r4i.log 2>&1; grep -E "^E  |^FAILED|passed|failed|Error" gpurun_out/r4i.log | cut -c1-300 | head -20
timeout 600 python bench.py --workload interactive --steps 10 --warmup 3 --no-cpu > gpurun_out/r4i_ia.json 2> gpurun_out/r4i_ia.err; tail -3 gpurun_out/r4i_ia.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r4i_ia.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value %.4g" % d["value"], "e2e", d["e2e"]["ms_per_step"], "dyn", d["config"]["momentum_thermo_ms_per_step"], "sort", d["config"]["sort_ms_per_step"], "err", d["config"]["device_error_flags"])
PY
