#!/bin/bash
mkdir -p gpurun_out
M=smsp__pipe_tensor_subpipe_dmma_cycles_active.avg,smsp__pipe_tensor_subpipe_dmma_cycles_active.max,smsp__pipe_tensor_subpipe_dmma_cycles_active.min,smsp__pipe_tensor_subpipe_dmma_cycles_active.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg,smsp__cycles_elapsed.max,gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,smsp__inst_executed_pipe_tensor_subpipe_dmma.max,smsp__inst_executed_pipe_tensor_subpipe_dmma.min
for nl in 100 96 80; do
timeout 600 ncu --metrics $M --clock-control none -k regex:cvscore_kernel -c 1 --csv --log-file gpurun_out/r2r_cv_$nl.csv python tools/bench_configs.py --configs 3 --reps 1 --scale 0.2 --nlambda $nl > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r2r_cv_$nl.csv') if not l.startswith('==')))
h=rows[0]; ni=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
print('--- nlambda $nl')
for r in rows[1:]:
    print(r[ni], r[vi], r[ui])
PY
done
