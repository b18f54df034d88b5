#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_mts_gpu.py::test_mts_with_runge_kutta_switches_to_verlet_like_the_reference tests/test_rk_interactions_gpu.py::test_rk4_with_footloose_is_refused -m gpu -q --no-header 2>&1 | tail -3
timeout 200 python scratch/ia_probe.py 1e6 10 2>&1 | tail -2
timeout 300 python scratch/ia_probe.py 1e7 5 2>&1 | tail -2
