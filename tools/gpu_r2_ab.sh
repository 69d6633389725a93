#!/bin/bash
# 2 GPUs: the 2-GPU bench line (incl. the e2e leg on a host that cannot pin 2 x 100 GB)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2ab_bench_n2.json 2> gpurun_out/r2ab_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2ab_bench_n2.json')); print(d['value'], d['e2e']); s=d['secondary']; print(s['logistic_configs3']['fit_s'], s['logistic_configs3']['phases_ms_rank0']); print(s['xval_configs2']['fit_s'], s['parity'])"; grep -v "^\s*$" gpurun_out/r2ab_bench_n2.err | tail -4 | cut -c1-300
