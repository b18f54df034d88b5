#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_interactions_gpu.py tests/test_footloose_gpu.py "tests/test_multirank_gpu.py::test_interacting_bergs_across_ranks" tests/test_trajectory.py -m gpu -q --no-header > gpurun_out/r4g.log 2>&1; grep -E "^E  |^FAILED|passed|failed|Error" gpurun_out/r4g.log | cut -c1-300 | head -20
timeout 200 python scratch/ia_probe.py 1e6 10 2>&1 | tail -1 | cut -c1-200
timeout 300 python scratch/ia_probe.py 1e7 5 2>&1 | tail -1 | cut -c1-200
KID_IA_NO_REC=1 timeout 200 python scratch/ia_probe.py 1e6 10 2>&1 | tail -1 | cut -c1-120
