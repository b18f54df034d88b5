#!/bin/bash
# round-2 second GPU pass: full parity suite, fast-path kernel variants, ncu of k_step_fast
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
B="python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-weak-base"
$B > gpurun_out/r2b_fast6.json 2> gpurun_out/r2b.err
KID_NO_FAST=1 $B > gpurun_out/r2b_nofast.json 2>> gpurun_out/r2b.err
for v in fm5 fm4 fm7; do KID_B200_LIB=$PWD/icebergs_b200/lib/var/libkid_$v.so $B > gpurun_out/r2b_$v.json 2>> gpurun_out/r2b.err; done
KID_SORT_INTERVAL=64 $B > gpurun_out/r2b_fast6_sort64.json 2>> gpurun_out/r2b.err
for f in fast6 nofast fm5 fm4 fm7 fast6_sort64; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_$f.json"))
    print("$f", "ms/step %.4f kern %.4f frac %.3f bergs %d sort/call %s launches %d" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["config"]["bergs_total"], d["config"]["sort_ms_per_call"], d["gpu_launches"]))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
tail -5 gpurun_out/r2b.err
ncu --set full --clock-control none --import-source on -k regex:k_step_fast -s 5 -c 1 -f -o gpurun_out/prof_kfast_r2b python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2b_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 33 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2b_ncu_launch.log 2>&1
