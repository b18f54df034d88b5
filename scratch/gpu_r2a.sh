#!/bin/bash
# round-2 first GPU pass: parity suite, bench (driver arguments and 50 steps), dense-tile sort timing, launch list, ncu of k_step
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2a_bench_n1_50.json 2>> gpurun_out/r2a_bench_n1.err
python bench.py --steps 32 --warmup 5 --no-cpu --no-e2e --no-weak-base --gni 360 --gnj 360 --bergs-per-gpu 12500000 > gpurun_out/r2a_bench_dense.json 2>> gpurun_out/r2a_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>> gpurun_out/r2a_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 33 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2a_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 5 -c 1 -f -o gpurun_out/prof_kstep_r2a python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-weak-base > gpurun_out/r2a_ncu_kstep.log 2>&1
for f in gpurun_out/r2a_bench_n1.json gpurun_out/r2a_bench_n1_50.json gpurun_out/r2a_bench_dense.json gpurun_out/r2a_bench_ref.json; do echo "== $f"; cut -c1-1500 $f; done
tail -5 gpurun_out/r2a_bench_n1.err
