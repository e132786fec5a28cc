#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort or radix or wide or stock_k or sorted or mixed" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -4 gpurun_out/pytest_new.log
python scripts/prof_sort.py; python scripts/prof_sort.py 4e6
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sort.csv python scripts/prof_sort.py > /dev/null 2>&1; python scripts/summarise_launches.py gpurun_out/launches_sort.csv | head -8
