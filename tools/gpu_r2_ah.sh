#!/bin/bash
# launch list of one logistic fit at the per-rank shard of an 8-GPU run (n = 2.5e5 x 1000): where do the 0.1 s go, kernel by kernel?
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2ah_launches_logistic_n250k.csv \
    python tools/bench_configs.py --configs 4 --scale 0.125 --reps 1 > gpurun_out/r2ah_ncu.log 2>&1; tail -1 gpurun_out/r2ah_ncu.log | cut -c1-300
timeout 120 python tools/bench_configs.py --configs 4 --scale 0.125 --reps 3 2>&1 | tail -1 | cut -c1-700
