#!/bin/bash
# 2 GPUs: whole GPU suite (incl. the 2-rank NCCL / peer-memory parity workers), then the 2-GPU bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2aa_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2aa_pytest_gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2aa_bench_n2.json 2> gpurun_out/r2aa_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2aa_bench_n2.json')); print(d['value'], d['e2e']['value']); s=d['secondary']; print(s['logistic_configs3']['fit_s'], s['logistic_configs3']['phases_ms_rank0']); print(s['xval_configs2']['fit_s'], s['parity'])"; tail -3 gpurun_out/r2aa_bench_n2.err | cut -c1-300
