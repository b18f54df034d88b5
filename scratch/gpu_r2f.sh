#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_mts_multirank_gpu.py -m gpu -q > gpurun_out/r2f_mts.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2f_mts.log | cut -c1-300
python -m pytest tests -m gpu -q --deselect tests/test_mts_multirank_gpu.py > gpurun_out/r2f_all.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2f_all.log | cut -c1-300
