#!/bin/bash
# round 2, call E: full GPU suite with the new entries, then ncu evidence for the logistic slab kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2e_pytest_gpu.log; grep -E "file-backed ingest|configs\[3\]-scale" gpurun_out/r2e_pytest_gpu.log
K='regex:logit_slab|ls_sum|slab_relayout|irls_|oem_path|clamp_one|gram_syrk|gram_reduce|colstats|vecsum|sum_partials|assemble'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file gpurun_out/r2e_launches_logistic.csv \
    python tools/bench_configs.py --configs 4 --reps 1 --scale 0.5 > gpurun_out/r2e_ncu_list.log 2>&1; tail -2 gpurun_out/r2e_ncu_list.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logit_slab_kernel -s 2 -c 1 -o gpurun_out/r2e_prof_logit_slab \
    python tools/bench_logit_pass.py --n 1000000 --reps 2 > gpurun_out/r2e_ncu_full.log 2>&1; tail -2 gpurun_out/r2e_ncu_full.log | cut -c1-300
ls -la gpurun_out | grep r2e
