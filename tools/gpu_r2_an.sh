#!/bin/bash
# NOTE: the advance-on-convergence variant these runs measured was dropped (no gain); OEMB200_IRLS_NO_ADVANCE no longer exists.
# 2 GPUs: multi-GPU tests, then the logistic leg with and without advance-on-convergence (same library, env switch)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -2
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 1 --warmup 3 --rows 1250000 --no-e2e --no-cpu --secondary-steps 4 2>/dev/null | grep '^{' | python -c "
import sys,json; d=json.loads(sys.stdin.read()); L=d['secondary']['logistic_configs3']; print('$1', L['fit_s'], L['phases_ms_rank0'], L['kernel_launches'], d['secondary']['parity']['max_dbeta_vs_n1'])"
}
run advance
OEMB200_IRLS_NO_ADVANCE=1 run no_advance
run advance
OEMB200_IRLS_NO_ADVANCE=1 run no_advance
