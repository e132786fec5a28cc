#!/bin/bash
# ncu --set full captures of the two dominant kernels + the RED-rate microbenchmark
set -x
mkdir -p gpurun_out
./scripts/microbench/red_rate > gpurun_out/red_rate.txt 2>&1; cat gpurun_out/red_rate.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:count_prefix -s 16 -c 2 -f -o gpurun_out/prof_count_prefix \
    python scripts/prof_count_all.py 2e7 > gpurun_out/prof1.log 2>&1; tail -2 gpurun_out/prof1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dedup_scan -s 1 -c 1 -f -o gpurun_out/prof_dedup_scan \
    python scripts/prof_count_all.py 2e7 > gpurun_out/prof2.log 2>&1; tail -2 gpurun_out/prof2.log
ls -la gpurun_out
