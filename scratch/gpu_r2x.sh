#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_rk_interactions_gpu.py tests/test_parity_gpu.py tests/test_interactions_gpu.py tests/test_footloose_gpu.py -m gpu -q > gpurun_out/r2x_test.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed" gpurun_out/r2x_test.log | cut -c1-300
