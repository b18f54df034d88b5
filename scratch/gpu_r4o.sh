#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_interactions_gpu.py -m gpu -q --no-header -k "crowd" > gpurun_out/r4o.log 2>&1; grep -E "^E  |^FAILED|passed|failed|Error" gpurun_out/r4o.log | cut -c1-400 | head -20
