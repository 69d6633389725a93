#!/bin/bash
# round 2, call G (8 GPUs): 8-rank parity of every row-sharded entry over the in-library communicator, then the 8-GPU bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2g_gpus.txt 2>&1; free -g | head -2 >> gpurun_out/r2g_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 tests/dist_gpu_worker.py > gpurun_out/r2g_dist8.log 2>&1; grep -E "DIST_OK|Error|error|assert" gpurun_out/r2g_dist8.log | head -5
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2g_bench_n8.json 2> gpurun_out/r2g_bench_n8.err; tail -c 6500 gpurun_out/r2g_bench_n8.json; tail -8 gpurun_out/r2g_bench_n8.err | cut -c1-300
