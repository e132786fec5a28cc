#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hamdist or workflow or consumers" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -6 gpurun_out/pytest_new.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hamdist_mma_kernel" -c 1 -o gpurun_out/prof_hamdist_mma -f python scripts/hamdist_bench.py 1e5 1 > gpurun_out/prof_hamdist_mma.log 2>&1; tail -2 gpurun_out/prof_hamdist_mma.log
( time python bench.py --no-workflow ) > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
