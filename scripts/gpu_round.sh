#!/bin/bash
# One GPU-box visit: parity tests, smoke, the default bench line, the reference arm, the ncu launch list of the same
# command, and one ncu --set full capture of the dominant kernels (on a smaller input: ncu replays every kernel ~40 times).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
( time python bench.py --extras ) > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; tail -c 1500 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/bench_under_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"partition_kernel|dedup_scan|bucket_hist|bucket_count" -s 4 -c 4 \
    -o gpurun_out/prof_count -f python scripts/prof_count_all.py 2e7 > gpurun_out/prof_count.log 2>&1
tail -2 gpurun_out/prof_count.log
