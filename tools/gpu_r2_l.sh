#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2l_pytest_gpu.log
timeout 600 python tools/bench_configs.py --configs 4 --reps 4 2>&1 | tail -1 | cut -c1-620
# round-2 launch list of the headline bench command (kernel share of a step), at 2e6 rows so that ncu finishes quickly
K='regex:gram_syrk|gram_reduce|oem_path|assemble|vecsum|sum_partials|colstats'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/r2l_launches_bench_rows2e6.csv \
    python bench.py --steps 2 --warmup 3 --rows 2000000 --no-e2e --no-cpu --no-secondary > gpurun_out/r2l_ncu_bench.log 2>&1; tail -1 gpurun_out/r2l_ncu_bench.log | cut -c1-200
