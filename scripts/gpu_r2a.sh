#!/bin/bash
# round-2 visit A: whole GPU suite, then the device-resident bench line with its phases
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 ) > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_quick.log') if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step']); print(d['roofline']['phases_ms']); print(d.get('hamdist')); print(d['checks'])
PY
