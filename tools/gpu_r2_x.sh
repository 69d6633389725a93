#!/bin/bash
# path kernel round: prof counters, timings without the counters, the whole GPU suite
mkdir -p gpurun_out
OEMB200_PATH_PROF=1 timeout 60 python tools/bench_sparse.py --n 1000000 --p 1000 --reps 1 2>&1 | grep -E "path prof" | tail -2 | cut -c1-420 | tee gpurun_out/r2x_path_prof.log
timeout 60 python tools/bench_sparse.py --n 1000000 --p 1000 --reps 3 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/r2x_sparse.json
OEMB200_PATH_DMMA=1 timeout 60 python tools/bench_sparse.py --n 1000000 --p 1000 --reps 3 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/r2x_sparse_dmma.json
timeout 120 python tools/bench_configs.py --configs 4 --reps 2 2>&1 | tail -1 | cut -c1-700 | tee gpurun_out/r2x_config4.json
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2x_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2x_pytest_gpu.log
