#!/bin/bash
# final 1-GPU evidence of the round: smoke, whole GPU suite, default bench line (timed by the shell as the driver would)
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2ad_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2ad_pytest_gpu.log
( time timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2ad_bench_n1.json 2> gpurun_out/r2ad_bench_n1.err ) 2>&1 | grep real
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2ad_bench_n1.json') if l.startswith('{')][-1]
print({k:d[k] for k in ('value','gpu_launches')}, d['config']['phases_ms'], d['e2e']['value'], d['e2e'].get('from_pageable',{}).get('value'), d['cpu_baseline']['value'])
s=d['secondary']; print(s['logistic_configs3']['fit_s'], s['logistic_configs3']['phases_ms_rank0']); print(s['xval_configs2']['fit_s'], s['xval_configs2']['phases_ms_rank0'])"; tail -2 gpurun_out/r2ad_bench_n1.err | cut -c1-300
