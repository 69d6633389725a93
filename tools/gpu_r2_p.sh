#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2p_pytest_gpu.log
timeout 900 python tools/bench_configs.py --configs 3,5 --reps 3 2>&1 | tail -2 | cut -c1-700
