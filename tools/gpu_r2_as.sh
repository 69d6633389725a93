#!/bin/bash
# slab route for small launch-bound logistic problems: whole GPU suite, the vignette config, configs[3] unchanged
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 120 python tools/bench_configs.py --configs 6 --reps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:75], round(d['wall_s']*1e3,2), 'ms')"
