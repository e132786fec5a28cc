#!/bin/bash
# GPU-box visit for the rows built after the counting path: parity suite, then the ingest / sort-path measurements
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
( time timeout 600 python scripts/bench_next.py ) > gpurun_out/next.log 2>&1; tail -c 3000 gpurun_out/next.log
