#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout -k 5 120 python scratch/debug_mr_long.py 2 1.5e6 $2 $1 > gpurun_out/r2t_$tag.log 2>&1; echo "== $tag: $(grep -c ' ok ' gpurun_out/r2t_$tag.log) stages ok; $(grep FAILED gpurun_out/r2t_$tag.log | head -1 | cut -c1-120)"; tail -1 gpurun_out/r2t_$tag.log | cut -c1-200; }
run 16x10 none
run 16x8 all
