#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 500 python -m pytest tests/test_fold_gpu.py -m gpu -q -x --no-header > gpurun_out/r4a_fold.log 2>&1; grep -E "^E  |^FAILED|passed|failed|Error" gpurun_out/r4a_fold.log | cut -c1-400 | head -40
