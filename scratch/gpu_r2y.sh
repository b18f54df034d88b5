#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_mts_gpu.py tests/test_mts_multirank_gpu.py tests/test_interactions_gpu.py -m gpu -q > gpurun_out/r2y_test.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2y_test.log | cut -c1-300
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2y_$1.json").read().strip().splitlines()[-1])
    print("$1", d["config"]["elements"], "ms/step %.3f dyn %.3f us/substep %.2f launches %d" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["us_per_substep"], d["gpu_launches"]))
except Exception as e: print("$1 FAILED", e)
PY
}
for n in 192 432 2320; do
  timeout 120 python bench.py --workload bonded --elements $n --steps 10 --warmup 3 --no-cpu > gpurun_out/r2y_cl_$n.json 2> gpurun_out/r2y_cl_$n.err; show cl_$n
done
KID_MTS_CLUSTER_MIN=1 timeout 120 python bench.py --workload bonded --elements 30 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2y_cl_30.json 2> gpurun_out/r2y_cl_30.err; show cl_30
timeout 120 python bench.py --workload bonded --elements 30 --steps 10 --warmup 3 > gpurun_out/r2y_one_30.json 2> gpurun_out/r2y_one_30.err; show one_30
python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_one_30.json").read().strip().splitlines()[-1]); print("cpu 30:", d["cpu_baseline"])
PY
