#!/bin/bash
mkdir -p gpurun_out
python scripts/hamdist_bench.py > gpurun_out/hamdist_bench.log 2>&1; cat gpurun_out/hamdist_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hamdist_mma_kernel|hamdist_planes" -c 2 -o gpurun_out/prof_hamdist -f python scripts/hamdist_bench.py 1e5 1 > gpurun_out/prof_hamdist.log 2>&1; tail -3 gpurun_out/prof_hamdist.log
