#!/bin/bash
# round 2, call I: the reference (CPU) arm at full size on the GPU box's host, then the full 1-GPU bench line
mkdir -p gpurun_out
free -g | head -2; nproc
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 3 --warmup 2 > gpurun_out/r2i_reference_n1.json 2> gpurun_out/r2i_reference_n1.err ) 2>&1 | grep real; cat gpurun_out/r2i_reference_n1.json | cut -c1-1500; tail -3 gpurun_out/r2i_reference_n1.err
timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_n1.json'))
print({k:d[k] for k in ('value','gpu_launches')}, d['e2e'], d['cpu_baseline'])
print(d['secondary']['logistic_configs3']); print(d['secondary']['xval_configs2']['fit_s'], d['secondary']['parity'])"; tail -3 gpurun_out/r2i_bench_n1.err
