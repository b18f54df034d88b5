#!/bin/bash
# usage: [BENCH_ARGS="..."] scratch/sweep_defs.sh "<defs1>" "<defs2>" ...  -> kernel ms for each build
for defs in "$@"; do
  KID_NVCC_EXTRA="$defs" python -m icebergs_b200.build --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e $BENCH_ARGS 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$defs] [$BENCH_ARGS]', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
