#!/bin/bash
# 2 GPUs, same box: secondary logistic leg with the current library (A), the one before the launch fusions (B), A again
mkdir -p gpurun_out
run() {
  OEMB200_LIB_PATH=$1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 1 --warmup 3 --rows 1250000 --no-e2e --no-cpu --secondary-steps 3 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); L=d['secondary']['logistic_configs3']; print('$2', L['fit_s'], L['phases_ms_rank0'], L['kernel_launches'], L.get('allreduce_avg_us'))"
}
run oem_b200/lib/liboem_b200.so A
run oem_b200/lib/liboem_b200_pathonly.so B
run oem_b200/lib/liboem_b200.so A
run oem_b200/lib/liboem_b200_pathonly.so B
