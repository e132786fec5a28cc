#!/bin/bash
# round-2 visit: whole GPU suite, smoke, default bench line, reference arm, ncu launch list of the bench command,
# ncu --set full of the four counting kernels (2e7 reads)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
( time python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; tail -c 600 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-hamdist --no-piece2 --no-workflow --no-crosscheck > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/bench_under_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"partition_kernel|dedup_scan|bucket_hist|bucket_count" -s 4 -c 4 \
    -o gpurun_out/prof_count -f python scripts/prof_count_all.py 2e7 > gpurun_out/prof_count.log 2>&1
tail -2 gpurun_out/prof_count.log
