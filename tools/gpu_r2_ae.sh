#!/bin/bash
# 1 GPU, same box: the full default bench (e2e legs included, no CPU leg) with the current library (A) and the one before the launch fusions (B)
mkdir -p gpurun_out
run() {
  OEMB200_LIB_PATH=$1 timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --secondary-steps 3 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); L=d['secondary']['logistic_configs3']; print('$2', d['value'], L['fit_s'], L['phases_ms_rank0'], L['kernel_launches'])"
}
run oem_b200/lib/liboem_b200.so A
run oem_b200/lib/liboem_b200_pathonly.so B
run oem_b200/lib/liboem_b200.so A
