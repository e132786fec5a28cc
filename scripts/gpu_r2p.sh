#!/bin/bash
# whole GPU suite + scan_motif at cfg3 scale (1e7 reads x 100 bp) with its ncu launch list
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
( time python scripts/bench_workflow.py 1e7 cfg3 ) > gpurun_out/workflow_cfg3.log 2>&1; head -c 1500 gpurun_out/workflow_cfg3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_scan_motif_cfg3.csv \
    python scripts/bench_workflow.py 1e7 cfg3 > gpurun_out/workflow_cfg3_under_ncu.log 2>&1
python scripts/summarise_launches.py gpurun_out/launches_scan_motif_cfg3.csv | head -40
