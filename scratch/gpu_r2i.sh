#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prefetch_gpu.py tests/test_restart_io.py -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2i_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "kern", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "unpipelined", d["e2e"]["unpipelined_ms_per_step"])
print("cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
PY
for n in 0 192 2320; do timeout 600 python bench.py --workload bonded --elements $n --steps 10 --warmup 3 > gpurun_out/r2i_bonded_$n.json 2> gpurun_out/r2i_bonded_$n.err; tail -3 gpurun_out/r2i_bonded_$n.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_bonded_$n.json").read().strip().splitlines()[-1])
print("bonded", d["config"]["elements"], "ms/step", d["ms_per_step"], "dyn ms", d["roofline"]["kernel_ms"], "us/substep", d["config"]["us_per_substep"], "launches", d["gpu_launches"], "cpu", d.get("cpu_baseline",{}).get("value"), "gpu", d["value"])
PY
done
