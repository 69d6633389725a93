#!/bin/bash
# One short gpurun call: smoke + the whole GPU suite, then the sparse entry's launch list (our kernels only) and one full
# ncu capture of its Gram kernel.  (tools/gpu_round.sh is the long evidence run for the dense headline path.)
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
K='regex:csc_|csr_|scan_|sparse_|sum_splits|sum_rows|row_pairs|assemble|scale_sym|oem_path|vecsum|sum_partials'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches_sparse.csv \
    python tools/bench_sparse.py --n 1000000 --p 1000 --reps 0 > gpurun_out/ncu_sparse_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sparse_gram -c 1 -o gpurun_out/prof_sparse_gram \
    python tools/bench_sparse.py --n 1000000 --p 1000 --reps 0 > gpurun_out/ncu_sparse_full.log 2>&1
ls -la gpurun_out | tail -6
