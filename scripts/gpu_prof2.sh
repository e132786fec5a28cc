#!/bin/bash
set -x
mkdir -p gpurun_out
./scripts/microbench/red_rate 2>&1 | grep smem > gpurun_out/red_rate2.txt; cat gpurun_out/red_rate2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:partition_kernel -s 1 -c 1 -f -o gpurun_out/prof_partition \
    python scripts/prof_count_all.py 2e7 > gpurun_out/prof3.log 2>&1; tail -2 gpurun_out/prof3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bucket_count -s 1 -c 1 -f -o gpurun_out/prof_bucket_count \
    python scripts/prof_count_all.py 2e7 > gpurun_out/prof4.log 2>&1; tail -2 gpurun_out/prof4.log
