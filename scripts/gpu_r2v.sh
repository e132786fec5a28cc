#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "sort or radix or wide or stock_k or sorted or mixed" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -4 gpurun_out/pytest_new.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5
python scripts/prof_sort.py
