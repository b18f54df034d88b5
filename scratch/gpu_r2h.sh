#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mts_multirank_gpu.py -m gpu -q > gpurun_out/r2h_mts.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2h_mts.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_mts_multirank_gpu.py > gpurun_out/r2h_all.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2h_all.log | cut -c1-300
