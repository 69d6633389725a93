#!/bin/bash
# round 2, call A: logistic routes + slab kernel parity, then the data-pass microbenchmark and configs[3]
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_entries.py -m gpu -q -x -k "logistic or logit" > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log
timeout 300 python tools/bench_logit_pass.py > gpurun_out/r2a_logit_pass.json 2> gpurun_out/r2a_logit_pass.err; cat gpurun_out/r2a_logit_pass.json; tail -3 gpurun_out/r2a_logit_pass.err
timeout 600 python tools/bench_configs.py --configs 4 --reps 2 > gpurun_out/r2a_config4.json 2> gpurun_out/r2a_config4.err; cat gpurun_out/r2a_config4.json; tail -3 gpurun_out/r2a_config4.err
OEMB200_LOGIT_ROUTE=sweeps timeout 600 python tools/bench_configs.py --configs 4 --reps 2 > gpurun_out/r2a_config4_sweeps.json 2>&1; cat gpurun_out/r2a_config4_sweeps.json
