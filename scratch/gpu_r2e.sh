#!/bin/bash
mkdir -p gpurun_out
python scratch/debug_mts_images.py > gpurun_out/r2e_debug.log 2>&1; tail -40 gpurun_out/r2e_debug.log | cut -c1-300
python -m pytest tests/test_mts_gpu.py -m gpu -q > gpurun_out/r2e_mts1.log 2>&1; tail -12 gpurun_out/r2e_mts1.log | cut -c1-300
python -m pytest tests/test_mts_multirank_gpu.py -m gpu -q > gpurun_out/r2e_mts.log 2>&1; tail -12 gpurun_out/r2e_mts.log | cut -c1-300
