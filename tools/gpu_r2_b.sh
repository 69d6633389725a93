#!/bin/bash
mkdir -p gpurun_out
for n in 100000 2000000; do
  echo "== n=$n"; timeout 300 python tools/bench_logit_pass.py --n $n --reps 10 2>&1 | tail -1 | cut -c1-700
  echo "== n=$n two CTAs/SM"; OEMB200_SLAB_CTAS=2 timeout 300 python tools/bench_logit_pass.py --n $n --reps 10 2>&1 | tail -1 | cut -c1-700
done
echo "== p=500"; timeout 300 python tools/bench_logit_pass.py --n 2000000 --p 500 --reps 10 2>&1 | tail -1 | cut -c1-700
echo "== p=500 two"; OEMB200_SLAB_CTAS=2 timeout 300 python tools/bench_logit_pass.py --n 2000000 --p 500 --reps 10 2>&1 | tail -1 | cut -c1-700
echo "== p=2000"; timeout 300 python tools/bench_logit_pass.py --n 1000000 --p 2000 --reps 10 2>&1 | tail -1 | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_entries.py -m gpu -q -x -k "logistic or logit" > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
OEMB200_SLAB_CTAS=2 timeout 600 python -m pytest tests/test_gpu_entries.py -m gpu -q -x -k "logistic or logit" 2>&1 | tail -3
timeout 600 python tools/bench_configs.py --configs 4 --reps 2 > gpurun_out/r2a_config4.json 2> gpurun_out/r2a_config4.err; cat gpurun_out/r2a_config4.json; tail -3 gpurun_out/r2a_config4.err
