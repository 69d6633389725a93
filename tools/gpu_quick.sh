#!/bin/bash
# One short gpurun call: headline bench line + the sparse entry's launch list and one full ncu capture of its Gram kernel.
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | cut -c1-700
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sparse.csv \
    python tools/bench_sparse.py --n 1000000 --p 1000 --reps 0 > gpurun_out/ncu_sparse_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sparse_gram -c 1 -o gpurun_out/prof_sparse_gram \
    python tools/bench_sparse.py --n 1000000 --p 1000 --reps 0 > gpurun_out/ncu_sparse_full.log 2>&1
ls -la gpurun_out | tail -8
