#!/bin/bash
mkdir -p gpurun_out
KID_B200_LIB=$PWD/icebergs_b200/lib/var/libkid_base13.so ncu --set full --clock-control none --import-source on -k regex:k_step_tma -s 5 -c 1 -f -o gpurun_out/prof_ktma_r2o python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2o_ncu.log 2>&1
tail -2 gpurun_out/r2o_ncu.log | cut -c1-200
