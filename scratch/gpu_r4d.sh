#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 500 python -m pytest tests/test_spreading_gpu.py tests/test_parity_gpu.py -m gpu -q --no-header > gpurun_out/r4d.log 2>&1; grep -E "^E  |^FAILED|passed|failed|Error" gpurun_out/r4d.log | cut -c1-400 | head -40
