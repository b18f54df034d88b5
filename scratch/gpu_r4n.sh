#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 110 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_fold_gpu.py -m gpu -q --no-header -k "static or forcing_halos or mass_spread or lx2-ly1 or 2-1" > gpurun_out/r4n_memcheck_fold.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds" gpurun_out/r4n_memcheck_fold.log | head -8
timeout -k 5 70 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_interactions_gpu.py -m gpu -q --no-header -k "first_steps" > gpurun_out/r4n_memcheck_ia.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds" gpurun_out/r4n_memcheck_ia.log | head -8
