#!/bin/bash
# usage: scratch/sweep_minblocks.sh "4 5 6 8"  -> kernel ms for each KID_MINBLOCKS build
for mb in $1; do
  KID_NVCC_EXTRA="-DKID_MINBLOCKS=$mb $2" python -m icebergs_b200.build --force > /dev/null 2>&1
  python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('minblocks=$mb', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
