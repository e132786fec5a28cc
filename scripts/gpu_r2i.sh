#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "onehot or streamed or hamdist" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -25 gpurun_out/pytest_new.log
python scripts/e2e_hostpack.py > gpurun_out/e2e_hostpack.log 2>&1; cat gpurun_out/e2e_hostpack.log
