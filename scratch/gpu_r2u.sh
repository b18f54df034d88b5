#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 100 python scratch/debug_uneven.py 2 5 > gpurun_out/r2u_a.log 2>&1; tail -22 gpurun_out/r2u_a.log | cut -c1-250
echo "--- no pipeline"; KID_NO_PIPELINE=1 timeout -k 5 100 python scratch/debug_uneven.py 2 5 > gpurun_out/r2u_b.log 2>&1; grep steps gpurun_out/r2u_b.log | cut -c1-200
echo "--- interval 32"; timeout -k 5 100 python scratch/debug_uneven.py 2 32 > gpurun_out/r2u_c.log 2>&1; grep steps gpurun_out/r2u_c.log | cut -c1-200
echo "--- 1 rank"; timeout -k 5 100 python scratch/debug_uneven.py 1 5 > gpurun_out/r2u_d.log 2>&1; grep steps gpurun_out/r2u_d.log | cut -c1-200
