#!/bin/bash
# One gpurun call: smoke, GPU tests, headline bench, ncu launch list + full captures of the dominant kernels.
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log | cut -c1-400
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-300
python tools/bench_configs.py --configs 1,2,3,4,5 > gpurun_out/bench_configs.log 2>&1; cut -c1-300 gpurun_out/bench_configs.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --rows 2000000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gram_syrk -s 1 -c 1 -o gpurun_out/prof_gram \
    python bench.py --rows 2000000 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:oem_path -s 1 -c 1 -o gpurun_out/prof_path \
    python bench.py --rows 2000000 --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none -k regex:colstats_kernel -s 4 -c 1 -o gpurun_out/prof_colstats \
    python tools/bench_configs.py --configs 4 --scale 0.5 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:cvscore_kernel -c 1 -o gpurun_out/prof_cvscore \
    python tools/bench_configs.py --configs 3 --scale 0.2 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:xb_kernel -s 2 -c 1 -o gpurun_out/prof_xb \
    python tools/bench_configs.py --configs 4 --scale 0.5 --reps 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_xval.csv \
    python tools/bench_configs.py --configs 3 --scale 0.2 --reps 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:oem_path_reg -c 1 -o gpurun_out/prof_pathreg \
    python tools/bench_configs.py --configs 2 --reps 1 > /dev/null 2>&1
OEMB200_PATH_PROF=1 python tools/bench_configs.py --configs 1,2 --reps 1 2>&1 | grep 'path prof' > gpurun_out/path_prof.log
# sparse entry: timings, launch list of our kernels, full capture of its Gram kernel; dense column-statistics A/B
python tools/bench_sparse.py --n 1000000 --p 1000 --density 0.01 > gpurun_out/bench_sparse.log 2>&1
python tools/bench_sparse.py --n 1000000 --p 1000 --density 0.05 --reps 2 >> gpurun_out/bench_sparse.log 2>&1
python tools/bench_dense_stats.py 2000000 512 > gpurun_out/bench_dense_stats.log 2>&1
K='regex:csc_|csr_|scan_|sparse_|sum_splits|sum_rows|row_pairs|assemble|scale_sym|oem_path|vecsum|sum_partials'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/launches_sparse.csv \
    python tools/bench_sparse.py --n 1000000 --p 1000 --reps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sparse_gram -c 1 -o gpurun_out/prof_sparse_gram \
    python tools/bench_sparse.py --n 1000000 --p 1000 --reps 0 > /dev/null 2>&1
ls -la gpurun_out
