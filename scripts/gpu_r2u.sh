#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "streamed or host_encoders or cfg" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -5 gpurun_out/pytest_new.log
python scripts/e2e_hostpack.py > gpurun_out/e2e_hostpack.log 2>&1; cat gpurun_out/e2e_hostpack.log
