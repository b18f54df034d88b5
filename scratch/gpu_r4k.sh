#!/bin/bash
mkdir -p gpurun_out
# launch list of the final build (drift workload): per-launch times are cold-cache and serialised, the shares are what counts
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4k_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r4k_ncu_launch.log 2>&1; tail -1 gpurun_out/r4k_ncu_launch.log | cut -c1-150
# launch list of one interactive step + full capture of k_ia_velocity at 10 M bergs
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r4k_ia_launches.csv python scratch/ia_probe.py 1e7 2 > gpurun_out/r4k_ia_launch.log 2>&1; tail -1 gpurun_out/r4k_ia_launch.log | cut -c1-150
ncu --set full --clock-control none --import-source on -k regex:k_ia_velocity -s 1 -c 1 -f -o gpurun_out/prof_kia_r4k python scratch/ia_probe.py 1e7 1 > gpurun_out/r4k_kia.log 2>&1; tail -1 gpurun_out/r4k_kia.log | cut -c1-150
ls -la gpurun_out/*.ncu-rep | tail -2
