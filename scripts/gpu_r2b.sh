#!/bin/bash
# round-2 visit B (2 GPUs): full-size config parity + the 2-rank NCCL tests
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_configs.py tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/pytest_cfg_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cfg_multi.log
tail -40 gpurun_out/pytest_cfg_multi.log
