#!/bin/bash
# path kernel (global mode): cycle counters only
OEMB200_PATH_PROF=1 timeout 60 python tools/bench_sparse.py --n 1000000 --p 1000 --reps 1 2>&1 | grep -E "path prof|ms_path" | tail -3 | cut -c1-420
