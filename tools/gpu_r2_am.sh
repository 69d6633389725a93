#!/bin/bash
# NOTE: the advance-on-convergence variant these runs measured was dropped (no gain); OEMB200_IRLS_NO_ADVANCE no longer exists.
# advance-on-convergence in the IRLS loop: logistic tests, then the 8-GPU-shard-size fit with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_entries.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -q -x -k "logistic or logit" 2>&1 | tail -2
for v in 0 1 0 1; do
  if [ $v = 1 ]; then export OEMB200_IRLS_NO_ADVANCE=1; else unset OEMB200_IRLS_NO_ADVANCE; fi
  echo "no_advance=$v"; timeout 200 python tools/bench_logit_small.py 2>&1 | head -1
done
