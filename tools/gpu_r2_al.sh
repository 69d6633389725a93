#!/bin/bash
# 8 GPUs, final code: 8-rank parity of every row-sharded entry, then the strong-scaling legs (primary at a reduced shard, no e2e / CPU legs)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 tests/dist_gpu_worker.py > gpurun_out/r2al_dist8.log 2>&1; grep -E "DIST_OK|Error|error|assert" gpurun_out/r2al_dist8.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29673 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e --no-cpu --rows 2000000 --secondary-steps 4 > gpurun_out/r2al_bench_n8.json 2> gpurun_out/r2al_bench_n8.err
python - <<PY
import json
txt=open('gpurun_out/r2al_bench_n8.json').read(); d=json.loads(txt[txt.index('{"metric"'):].splitlines()[0])
s=d['secondary']; l=s['logistic_configs3']; print('logistic', l['fit_s'], l['phases_ms_rank0'], 'ar_us', l['allreduce_avg_us'], 'launches', l['kernel_launches'])
x=s['xval_configs2']; print('xval', x['fit_s'], x['phases_ms_rank0'], 'parity', s['parity']['max_dbeta_vs_n1'], s['allreduce_probe'])
PY
tail -3 gpurun_out/r2al_bench_n8.err | cut -c1-300
