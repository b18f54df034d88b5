#!/bin/bash
# round-2 fourth GPU pass: MTS across ranks, full parity suite, bench with the final fast kernel
mkdir -p gpurun_out
python -m pytest tests/test_mts_multirank_gpu.py -m gpu -q -x > gpurun_out/r2d_mts.log 2>&1; echo "mts pytest rc=$?"
tail -25 gpurun_out/r2d_mts.log | cut -c1-400
python -m pytest tests -m gpu -q --deselect tests/test_mts_multirank_gpu.py > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2d_pytest.log | cut -c1-400
B="python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-weak-base"
$B > gpurun_out/r2d_p5.json 2> gpurun_out/r2d.err
KID_NO_FAST=1 $B > gpurun_out/r2d_nofast.json 2>> gpurun_out/r2d.err
for f in p5 nofast; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_$f.json"))
    print("$f", "ms/step %.4f kern %.4f frac %.3f bergs %d sort/call %s launches %d slow %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["config"]["bergs_total"], d["config"]["sort_ms_per_call"], d["gpu_launches"], d["roofline"].get("slow_list_fraction")))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
tail -3 gpurun_out/r2d.err
