#!/bin/bash
# round 2, call N (8 GPUs): secondary legs with and without the speculative IRLS pipeline (primary at a reduced shard: not the subject here)
mkdir -p gpurun_out
for mode in spec nospec; do
  if [ $mode = nospec ]; then export OEMB200_IRLS_NO_SPECULATION=1; else unset OEMB200_IRLS_NO_SPECULATION; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2967$([ $mode = spec ] && echo 3 || echo 4) bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e --rows 2000000 --secondary-steps 4 > gpurun_out/r2n_bench_n8_$mode.json 2> gpurun_out/r2n_bench_n8_$mode.err
  python - <<PY
import json
txt=open('gpurun_out/r2n_bench_n8_$mode.json').read(); d=json.loads(txt[txt.index('{"metric"'):].splitlines()[0])
s=d['secondary']; l=s['logistic_configs3']; print('$mode', 'logistic', l['fit_s'], l['phases_ms_rank0'], 'ar_us', l['allreduce_avg_us'], 'launches', l['kernel_launches'])
print('$mode', 'xval', s['xval_configs2']['fit_s'], 'parity', s['parity']['max_dbeta_vs_n1'], s['allreduce_probe'])
PY
done
