#!/bin/bash
# path kernel (global mode) iteration experiments: cycle counters of CTA 0 + the two workloads the phase matters for
mkdir -p gpurun_out
OEMB200_PATH_PROF=1 timeout 60 python tools/bench_sparse.py --n 1000000 --p 1000 --reps 2 2>&1 | grep -E "path prof|ms_path" | tail -3 | cut -c1-400
timeout 120 python tools/bench_configs.py --configs 4 --reps 2 2>&1 | tail -1 | cut -c1-700
timeout 300 python -m pytest tests/test_gpu_entries.py tests/test_gpu_sparse.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3
