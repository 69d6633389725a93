#!/bin/bash
# round 2, call C: full GPU suite, then the 1-GPU bench line (primary + e2e + secondary + cpu baseline)
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2c_pytest_gpu.log; grep "file-backed ingest" gpurun_out/r2c_pytest_gpu.log
timeout 1200 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 6000 gpurun_out/r2c_bench_n1.json; tail -5 gpurun_out/r2c_bench_n1.err
