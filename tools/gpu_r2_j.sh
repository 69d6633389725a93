#!/bin/bash
# round 2, call J (2 GPUs): speculative IRLS pipeline -- single-GPU logistic tests, 2-rank parity, then the 2-GPU bench (secondary only matters)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "logistic or logit or multi or libcomm" > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
timeout 600 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | cut -c1-600
OEMB200_IRLS_NO_SPECULATION=1 timeout 600 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | cut -c1-600
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_n2.json')); print(d['value']); s=d['secondary']; print(s['logistic_configs3']); print(s['xval_configs2']['fit_s'], s['parity']['max_dbeta_vs_n1'])"; tail -3 gpurun_out/r2j_bench_n2.err | cut -c1-300
