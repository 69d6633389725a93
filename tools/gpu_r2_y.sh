#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path_modes.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2y_pytest_path_modes.log
timeout 250 compute-sanitizer --tool memcheck python tools/sanitize_path.py > gpurun_out/r2y_memcheck_path.log 2>&1; tail -4 gpurun_out/r2y_memcheck_path.log
timeout 250 compute-sanitizer --tool racecheck python tools/sanitize_path.py > gpurun_out/r2y_racecheck_path.log 2>&1; tail -4 gpurun_out/r2y_racecheck_path.log
