#!/bin/bash
mkdir -p gpurun_out
for mode in spec nospec; do
  if [ $mode = nospec ]; then export OEMB200_IRLS_NO_SPECULATION=1; else unset OEMB200_IRLS_NO_SPECULATION; fi
  timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --rows 2000000 --secondary-steps 4 > gpurun_out/r2k_$mode.json 2>gpurun_out/r2k_$mode.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/r2k_$mode.json')); s=d['secondary']['logistic_configs3']; print('$mode', s['fit_s'], s['phases_ms_rank0'], s['kernel_launches'])"
done
for mode in spec nospec; do
  if [ $mode = nospec ]; then export OEMB200_IRLS_NO_SPECULATION=1; else unset OEMB200_IRLS_NO_SPECULATION; fi
  timeout 600 python tools/bench_configs.py --configs 4 --reps 4 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$mode configs', d['wall_s'], d['phases_ms']['ms_total'])"
done
