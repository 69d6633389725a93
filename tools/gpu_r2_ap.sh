#!/bin/bash
# register-resident path kernel (q <= 256) with the branch-free precomputed prox: README configs + cycle counters + tests
OEMB200_PATH_PROF=1 timeout 120 python tools/bench_configs.py --configs 1,2 --reps 2 2>&1 | grep -E "register variant" | tail -2 | cut -c1-200
timeout 120 python tools/bench_configs.py --configs 1,2 --reps 3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:40], round(d['wall_s']*1e3,3), 'ms  path', d['phases_ms']['ms_path'])"
timeout 600 python -m pytest tests/test_gpu_entries.py tests/test_gpu_fuzz.py tests/test_reference_pins.py -m gpu -q -x 2>&1 | tail -2
