#!/bin/bash
# DRAM traffic of the Gram kernel at the headline shard size (one launch), for bench.py's roofline.traffic.
set -x
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gram_syrk -s 2 -c 1 --csv \
    --log-file gpurun_out/gram_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/gram_traffic.log 2>&1
grep gram_syrk gpurun_out/gram_traffic.csv | cut -d, -f12-
