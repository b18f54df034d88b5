#!/bin/bash
# round-2 third GPU pass: parity suite with the warp-cached fast kernel, variants, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
B="python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-weak-base"
$B > gpurun_out/r2c_c8m5.json 2> gpurun_out/r2c.err
KID_NO_FAST=1 $B > gpurun_out/r2c_nofast.json 2>> gpurun_out/r2c.err
for v in c8m6 c8m4 c6m5 c12m5; do KID_B200_LIB=$PWD/icebergs_b200/lib/var/libkid_$v.so $B > gpurun_out/r2c_$v.json 2>> gpurun_out/r2c.err; done
for f in c8m5 nofast c8m6 c8m4 c6m5 c12m5; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_$f.json"))
    print("$f", "ms/step %.4f kern %.4f frac %.3f bergs %d sort/call %s launches %d slow %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["config"]["bergs_total"], d["config"]["sort_ms_per_call"], d["gpu_launches"], d["roofline"].get("slow_list_fraction")))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
tail -5 gpurun_out/r2c.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_full.err; cut -c1-3000 gpurun_out/r2c_bench_full.json
ncu --set full --clock-control none --import-source on -k regex:k_step_fast -s 5 -c 1 -f -o gpurun_out/prof_kfast_r2c python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2c_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 33 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2c_ncu_launch.log 2>&1
