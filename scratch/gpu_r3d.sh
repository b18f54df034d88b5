#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3d_launches.csv python bench.py --steps 33 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r3d_ncu_launch.log 2>&1
tail -1 gpurun_out/r3d_ncu_launch.log | cut -c1-200
