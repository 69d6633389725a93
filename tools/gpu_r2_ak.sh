#!/bin/bash
# 2 GPUs: multi-GPU tests with the final library, then same-box A/B of the logistic leg (A = current, B = before the IRLS epilogue / device clocks)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -2
run() {
  OEMB200_LIB_PATH=$1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 1 --warmup 3 --rows 1250000 --no-e2e --no-cpu --secondary-steps 3 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); L=d['secondary']['logistic_configs3']; print('$2', L['fit_s'], L['phases_ms_rank0'], L['kernel_launches'], L.get('allreduce_avg_us'), d['secondary']['parity']['max_dbeta_vs_n1'])"
}
run oem_b200/lib/liboem_b200.so A
run oem_b200/lib/liboem_b200_prev.so B
run oem_b200/lib/liboem_b200.so A
