#!/bin/bash
# round-2 final evidence: launch list of the headline bench command (kernel share of a step) and one full ncu capture of the
# reworked path kernel (global mode, register mat-vec) on the sparse-entry workload
mkdir -p gpurun_out
K='regex:gram_syrk|gram_reduce|oem_path|assemble|vecsum|sum_partials|colstats'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/r2af_launches_bench_rows2e6.csv \
    python bench.py --steps 2 --warmup 3 --rows 2000000 --no-e2e --no-cpu --no-secondary > gpurun_out/r2af_ncu_bench.log 2>&1; tail -1 gpurun_out/r2af_ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oem_path_kernel -s 1 -c 1 -o gpurun_out/r2af_prof_path \
    python tools/bench_sparse.py --n 1000000 --p 1000 --reps 1 > gpurun_out/r2af_ncu_path.log 2>&1; tail -2 gpurun_out/r2af_ncu_path.log | cut -c1-200
ls -la gpurun_out/r2af*
