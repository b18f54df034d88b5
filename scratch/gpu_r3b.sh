#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 500 python -m pytest tests/test_mts_gpu.py -m gpu -q --durations=6 > gpurun_out/r3b_test.log 2>&1; grep -E "^E  .*(Error|assert|Fatal)|^FAILED|passed|failed|s call" gpurun_out/r3b_test.log | cut -c1-200
KID_MTS_NO_SMEM=1 timeout -k 5 300 python -m pytest tests/test_mts_gpu.py -m gpu -q -k "beam" --durations=3 > gpurun_out/r3b_test2.log 2>&1; grep -E "^FAILED|passed|failed|s call" gpurun_out/r3b_test2.log | cut -c1-200
