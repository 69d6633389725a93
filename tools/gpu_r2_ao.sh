#!/bin/bash
# NOTE: the OEMB200_FG_COLS switch this run used was an experiment (4 columns per CTA stayed); it no longer exists.
# fold gather: 4 / 8 / 16 columns per CTA (index traffic 12.5 / 6 / 3 % of the matrix) on configs[2], then the xval tests
for c in 4 8 16 4 8 16; do
  OEMB200_FG_COLS=$c timeout 120 python tools/bench_configs.py --configs 3 --reps 3 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms']; print('cols $c wall', round(d['wall_s']*1e3,1), 'total', round(p['ms_total'],1), 'gram', round(p['ms_gram'],1), 'cv', round(p['ms_cvscore'],1), 'path', round(p['ms_path'],1), 'gather+gaps', round(p['ms_total']-p['ms_gram']-p['ms_cvscore']-p['ms_path']-p['ms_h2d']-p['ms_colstats']-p['ms_assemble'],1))"
done
timeout 300 python -m pytest tests/test_gpu_entries.py tests/test_gpu_fullsize.py -m gpu -q -x -k "xval" 2>&1 | tail -2
