#!/bin/bash
# One gpurun call: smoke, small bench (sanity), full bench, ncu launch list + full capture of the Gram kernel.
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py --rows 1000000 --steps 2 --warmup 3 --cpu-sample-rows 50000 > gpurun_out/bench_small.log 2>&1; tail -2 gpurun_out/bench_small.log
python bench.py > gpurun_out/bench_full.log 2>&1; tail -2 gpurun_out/bench_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --rows 2000000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_syrk -s 1 -c 1 -o gpurun_out/prof_gram \
    python bench.py --rows 2000000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
