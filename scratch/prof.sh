#!/bin/bash
# usage (on the GPU box): scratch/prof.sh <tag>   -> gpurun_out/prof_kstep_<tag>.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:k_step -s 5 -c 1 -f -o gpurun_out/prof_kstep_$1 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log | cut -c1-200
