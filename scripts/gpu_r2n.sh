#!/bin/bash
# 2-GPU visit: NCCL tests (sharded count_kmers, scattered merge, scan_motif under torchrun) + the 2-GPU bench line
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -30 gpurun_out/pytest_multi.log
bash scripts/gpu_r2m.sh 2
