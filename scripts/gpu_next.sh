#!/bin/bash
# GPU-box visit for the rows built after the counting path: parity suite, the ingest / sort-path measurements, and an
# ncu --set full capture of their kernels (smaller input: ncu replays every kernel ~40 times)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
( time timeout 600 python scripts/bench_next.py ) > gpurun_out/next.log 2>&1; tail -c 3000 gpurun_out/next.log
if [ "$1" = "prof" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fasta_tile|fasta_emit" -c 3 \
    -o gpurun_out/prof_fasta -f python scripts/bench_next.py 1e6 5e5 prof > gpurun_out/prof_fasta.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"onesweep_|dedup_keys|window_keys|head_|merge_plan|merge_write|run_length" -c 45 \
    -o gpurun_out/prof_next -f python scripts/bench_next.py 1e6 5e5 prof > gpurun_out/prof_next.log 2>&1
tail -2 gpurun_out/prof_next.log
fi
