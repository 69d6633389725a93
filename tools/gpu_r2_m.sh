#!/bin/bash
for mode in spec nospec; do
  if [ $mode = nospec ]; then export OEMB200_IRLS_NO_SPECULATION=1; else unset OEMB200_IRLS_NO_SPECULATION; fi
  echo "== $mode"
  OEMB200_TIMING=1 timeout 600 python tools/bench_configs.py --configs 4 --reps 2 2>&1 | grep -E "timing|wall_s" | tail -9 | cut -c1-200
done
