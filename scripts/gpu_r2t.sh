#!/bin/bash
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py 3000 ) > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
tail -12 gpurun_out/memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py 1500 ) > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/racecheck.log
tail -8 gpurun_out/racecheck.log
