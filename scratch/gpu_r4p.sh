#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_fold_gpu.py -m gpu -q --no-header -k "nccl" > gpurun_out/r4p.log 2>&1; grep -E "^E  |^FAILED|passed|failed|Error|skipped" gpurun_out/r4p.log | cut -c1-400 | head -10
