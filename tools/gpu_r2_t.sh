#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2t_bench_n1.json'))
print(d['value'], d['e2e']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
s=d['secondary']; print(s['logistic_configs3']['fit_s'], s['logistic_configs3']['phases_ms_rank0']); print(s['xval_configs2']['fit_s'], s['xval_configs2']['phases_ms_rank0'])"; tail -3 gpurun_out/r2t_bench_n1.err
