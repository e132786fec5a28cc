#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "streamed or host_encoders or cfg" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -5 gpurun_out/pytest_new.log
( time python scripts/e2e_hostpack.py ) > gpurun_out/e2e_hostpack.log 2>&1; tail -12 gpurun_out/e2e_hostpack.log
