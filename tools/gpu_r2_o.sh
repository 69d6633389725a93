#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2o_pytest_gpu.log
