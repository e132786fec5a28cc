#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python scripts/sweep_part.py 1e8 > gpurun_out/sweep_part.txt 2>&1; cat gpurun_out/sweep_part.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_part.csv \
    python scripts/prof_count_all.py 1e8 > gpurun_out/prof_part.log 2>&1; tail -2 gpurun_out/prof_part.log
