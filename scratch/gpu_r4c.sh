#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-weak-base > gpurun_out/r4c_$name.json 2> gpurun_out/r4c_$name.err
  grep "KID_L2_PERSIST" gpurun_out/r4c_$name.err | head -1
  python - <<PY
import json
d=json.loads(open("gpurun_out/r4c_$name.json").read().strip().splitlines()[-1])
print("$name", "ms/step", round(d["ms_per_step"],4), "kern", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],4), d["clocks"]["reasons"], d["clocks"]["sm_mhz"])
PY
}
run base A=1
run persist KID_L2_PERSIST=1
run stream KID_B200_LIB=$PWD/icebergs_b200/lib/libkid_stream.so
run stream_persist KID_B200_LIB=$PWD/icebergs_b200/lib/libkid_stream.so KID_L2_PERSIST=1
run base2 A=1
