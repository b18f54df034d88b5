#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scratch/debug_r2g.py B > gpurun_out/r2g_debug.log 2>&1; tail -30 gpurun_out/r2g_debug.log | cut -c1-260
timeout 900 python -m pytest tests/test_mts_multirank_gpu.py tests/test_mts_gpu.py -m gpu -q > gpurun_out/r2g_mts.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2g_mts.log | cut -c1-300
