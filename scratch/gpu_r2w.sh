#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_trajectory.py tests/test_multirank_gpu.py -m gpu -q > gpurun_out/r2w_test.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2w_test.log | cut -c1-300
