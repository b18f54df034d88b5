#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 400 python -m pytest tests/test_fold_gpu.py tests/test_mts_gpu.py::test_mts_with_runge_kutta_switches_to_verlet_like_the_reference tests/test_rk_interactions_gpu.py::test_rk4_with_footloose_is_refused "tests/test_multirank_gpu.py::test_nccl_ranks_match_single_rank_oracle" -m gpu -q --no-header > gpurun_out/r4e.log 2>&1; grep -E "^E  |^FAILED|passed|failed|Error|skipped" gpurun_out/r4e.log | cut -c1-400 | head -40
