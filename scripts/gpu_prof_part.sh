#!/bin/bash
# ncu --set full of the level-kmax count kernels (bucket_hist, partition, bucket_count) on 2e7 reads x 100 bp
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"partition_kernel|bucket_hist|bucket_count" -s 3 -c 3 \
    -o gpurun_out/prof_part -f python scripts/prof_count_all.py 2e7 > gpurun_out/prof_part.log 2>&1
tail -2 gpurun_out/prof_part.log
