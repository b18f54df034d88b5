#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_step_fast -s 5 -c 1 -f -o gpurun_out/prof_kfast_r2k python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2k_ncu.log 2>&1
tail -2 gpurun_out/r2k_ncu.log | cut -c1-200
