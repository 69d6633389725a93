#!/bin/bash
# slab route below p = 128?  vignette logistic (p = 100) with the two-sweep route (default) and the slab route, plus the logistic tests under the switch
for m in 128 64; do
  OEMB200_SLAB_MIN_P=$m timeout 120 python tools/bench_configs.py --configs 6 --reps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); p=d['phases_ms']; print('min_p $m', d['config'][30:75], round(d['wall_s']*1e3,2), 'ms  xb', p['ms_irls_xb'], 'xtr', p['ms_irls_xtr'], 'path', p['ms_path'], 'launches', d['kernel_launches'])"
done
OEMB200_SLAB_MIN_P=16 timeout 600 python -m pytest tests/test_gpu_entries.py tests/test_gpu_fuzz.py -m gpu -q -x -k "logistic or logit" 2>&1 | tail -2
