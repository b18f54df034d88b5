#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload interactive --steps 10 --warmup 3 > gpurun_out/r4h_ia.json 2> gpurun_out/r4h_ia.err; tail -3 gpurun_out/r4h_ia.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r4h_ia.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value %.4g" % d["value"], "e2e", d["e2e"]["ms_per_step"], "dyn", d["config"]["momentum_thermo_ms_per_step"], "sort", d["config"]["sort_ms_per_step"], "cand", d["roofline"]["candidates_per_berg"], "frac", d["roofline"]["frac"], "cpu %.4g" % d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "err", d["config"]["device_error_flags"], "launches", d["gpu_launches"], d["clocks"]["reasons"])
PY
