#!/bin/bash
mkdir -p gpurun_out
B="timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-weak-base"
show() { python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2m_$1.json").read().strip().splitlines()[-1])
    print("$1", "ms/step %.4f kern %.4f frac %.3f launches %d clocks %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$1", "FAILED", e)
PY
}
for v in c17m5 c13m5 c13m6 c13m7; do KID_B200_LIB=$PWD/icebergs_b200/lib/var/libkid_$v.so $B > gpurun_out/r2m_$v.json 2>> gpurun_out/r2m.err; show $v; done
KID_NO_TMA=1 $B > gpurun_out/r2m_notma.json 2>> gpurun_out/r2m.err; show notma
for v in c13m6 c13m7; do KID_B200_LIB=$PWD/icebergs_b200/lib/var/libkid_$v.so $B > gpurun_out/r2m_$v.json 2>> gpurun_out/r2m.err; show $v; done
tail -5 gpurun_out/r2m.err
