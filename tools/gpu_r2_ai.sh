#!/bin/bash
# device-side phase clocks in the IRLS loop: logistic tests, then the small-shard fit (wall / phases) and configs[3]
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_entries.py tests/test_gpu_fullsize.py -m gpu -q -x -k "logistic or logit" 2>&1 | tail -2
timeout 200 python tools/bench_logit_small.py 2>&1 | tail -4
timeout 120 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | cut -c1-600
