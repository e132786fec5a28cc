#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "onehot or hamdist" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
timeout 120 python scripts/hamdist_bench.py > gpurun_out/hamdist_bench.log 2>&1; cat gpurun_out/hamdist_bench.log
KMAP_HAMDIST_MMA_NO_TMA=1 timeout 120 python scripts/hamdist_bench.py 2>&1 | grep GEMM
