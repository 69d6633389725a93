#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:oem_path_kernel -c 1 -o gpurun_out/r2s_prof_path python tools/bench_sparse.py --n 1000000 --p 1000 --reps 0 > gpurun_out/r2s_ncu.log 2>&1; tail -2 gpurun_out/r2s_ncu.log | cut -c1-200; ls -la gpurun_out/r2s_prof_path.ncu-rep
