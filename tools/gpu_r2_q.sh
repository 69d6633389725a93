#!/bin/bash
for nl in 100 96 80 107; do
  timeout 600 python tools/bench_configs.py --configs 3 --reps 2 --scale 0.4 --nlambda $nl 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($nl*3, 'cols  ms_cvscore', d['phases_ms']['ms_cvscore'], 'tflops', round(d['cvscore_tflops'],2))"
done
