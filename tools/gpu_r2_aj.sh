#!/bin/bash
# same-box A/B at full size (configs[3], one GPU): current library (A) vs the committed one before the device clocks / IRLS epilogue (B)
run() {
  OEMB200_LIB_PATH=$1 timeout 120 python tools/bench_configs.py --configs 4 --reps 3 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); p=d['phases_ms']; print('$2 wall', round(d['wall_s']*1e3,1), 'total', round(p['ms_total'],1), 'xb', round(p['ms_irls_xb'],1), 'path', round(p['ms_path'],1), 'gram', round(p['ms_gram'],1), 'rest', round(p['ms_total']-p['ms_irls_xb']-p['ms_path']-p['ms_gram']-p['ms_colstats']-p['ms_relayout'],1), 'launches', d['kernel_launches'])"
}
run oem_b200/lib/liboem_b200.so A
run oem_b200/lib/liboem_b200_prev.so B
run oem_b200/lib/liboem_b200.so A
run oem_b200/lib/liboem_b200_prev.so B
