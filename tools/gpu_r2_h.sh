#!/bin/bash
mkdir -p gpurun_out
for cpc in 8 16; do
echo "== min cpc $cpc"
OEMB200_PATH_MIN_CPC=$cpc OEMB200_PATH_PROF=1 timeout 300 python tools/bench_sparse.py --n 1000000 --p 1000 --reps 2 2>&1 | grep -E "path prof\] mode|ms_path" | tail -2 | cut -c1-330
done
timeout 600 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | cut -c1-500
OEMB200_PATH_MIN_CPC=16 timeout 600 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | cut -c1-500
