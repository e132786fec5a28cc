#!/bin/bash
# counting-path tests + device-resident bench line (no e2e / cpu legs)
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "count_all or cfg3 or streamed or large_input or dedup or partitioned" ) > gpurun_out/pytest_quick.log 2>&1; tail -4 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --no-e2e --no-cpu --no-hamdist --steps 5 --warmup 3 $1 ) > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_quick.log') if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step']); print(d['roofline']['phases_ms']); print(d['checks'])
PY
