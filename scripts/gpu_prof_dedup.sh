#!/bin/bash
# ncu --set full of the per-read scan (dedup_scan_kernel) on 2e7 reads x 100 bp
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dedup_scan" -s 1 -c 1 \
    -o gpurun_out/prof_dedup -f python scripts/prof_count_all.py 2e7 > gpurun_out/prof_dedup.log 2>&1
tail -2 gpurun_out/prof_dedup.log
