mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 100 python tools/bench_dense_stats.py 2000000 512 > gpurun_out/bench_dense_stats.log 2>&1; cut -c1-500 gpurun_out/bench_dense_stats.log
