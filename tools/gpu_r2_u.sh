#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2u_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2u_pytest_gpu.log
timeout 600 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | cut -c1-600
