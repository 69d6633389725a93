#!/bin/bash
# generic path kernel (one CTA / cluster / DMMA global modes) with the branch-free prox in the mat-vec epilogue: xval + small logistic + tests
timeout 120 python tools/bench_configs.py --configs 3,6 --reps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:60], round(d['wall_s']*1e3,2), 'ms  path', d['phases_ms']['ms_path'])"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
