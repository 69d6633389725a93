#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 120 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms']; print('timers on  wall', round(d['wall_s']*1e3,1), 'total', p['ms_total'], 'xb', p['ms_irls_xb'], 'gaps', round(p['ms_total']-sum(v for k,v in p.items() if k!='ms_total'),1))"
OEMB200_NO_PHASE_TIMERS=1 timeout 120 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('timers off wall', round(d['wall_s']*1e3,1))"
done
