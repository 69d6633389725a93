#!/bin/bash
# round 2, call D (2 GPUs): row-sharded parity over both all-reduce transports, then the 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s > gpurun_out/r2d_pytest_multi.log 2>&1; tail -5 gpurun_out/r2d_pytest_multi.log; grep DIST_OK gpurun_out/r2d_pytest_multi.log
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; tail -c 5000 gpurun_out/r2d_bench_n2.json; tail -8 gpurun_out/r2d_bench_n2.err
