#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hamdist_mma_kernel" -c 1 -o gpurun_out/prof_hamdist_mma -f python scripts/hamdist_bench.py 1e5 1 > gpurun_out/prof_hamdist_mma.log 2>&1; tail -2 gpurun_out/prof_hamdist_mma.log
( time python scripts/bench_next.py ) > gpurun_out/next.log 2>&1; tail -c 1500 gpurun_out/next.log
( time python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
