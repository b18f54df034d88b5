#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q > gpurun_out/r4j_pytest.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r4j_pytest.log | cut -c1-300 | head -30
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r4j_bench.json 2> gpurun_out/r4j_bench.err; tail -2 gpurun_out/r4j_bench.err | cut -c1-200; python - <<PY
import json
d=json.loads(open("gpurun_out/r4j_bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value %.4g" % d["value"], "kern", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "unpipelined", d["e2e"]["unpipelined_ms_per_step"], "weak", d["weak_base"]["ms_per_step"], d["weak_base"]["kernel_ms"], "cpu %.4g" % d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"]["reasons"])
PY
timeout 600 python bench.py --workload interactive --steps 10 --warmup 3 > gpurun_out/r4j_ia.json 2> gpurun_out/r4j_ia.err; tail -2 gpurun_out/r4j_ia.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r4j_ia.json").read().strip().splitlines()[-1])
print("IA ms/step", d["ms_per_step"], "value %.4g" % d["value"], "e2e", d["e2e"]["ms_per_step"], "dyn", d["config"]["momentum_thermo_ms_per_step"], "sort", d["config"]["sort_ms_per_step"], "cpu %.4g" % d["cpu_baseline"]["value"], "err", d["config"]["device_error_flags"])
PY
