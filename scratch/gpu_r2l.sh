#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_multirank_gpu.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2l_pytest.log | cut -c1-300
B="timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-weak-base"
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_$1.json").read().strip().splitlines()[-1])
    print("$1", "ms/step %.4f kern %.4f frac %.3f sort/call %s launches %d slow %s clocks %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["config"]["sort_ms_per_call"], d["gpu_launches"], d["roofline"].get("slow_list_fraction"), d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$1", "FAILED", e)
PY
}
$B > gpurun_out/r2l_tma.json 2> gpurun_out/r2l.err; show tma
KID_NO_TMA=1 $B > gpurun_out/r2l_notma.json 2>> gpurun_out/r2l.err; show notma
KID_TMA_CTAS=4 $B > gpurun_out/r2l_tma4.json 2>> gpurun_out/r2l.err; show tma4
KID_TMA_CTAS=3 $B > gpurun_out/r2l_tma3.json 2>> gpurun_out/r2l.err; show tma3
for v in pf0 pf1; do KID_B200_LIB=$PWD/icebergs_b200/lib/var/libkid_$v.so $B > gpurun_out/r2l_$v.json 2>> gpurun_out/r2l.err; show $v; done
tail -5 gpurun_out/r2l.err
