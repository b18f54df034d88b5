#!/bin/bash
# usage: scratch/sweep_env.sh "VAR=val ..." ...   -> bench line per environment
for e in "$@"; do
  env $e python bench.py --steps 32 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('[$e]', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
