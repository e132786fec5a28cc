#!/bin/bash
# Final 1-GPU visit of round 2: the GPU suite with and without the folded run-end corrections (KMAP_FOLD_RUN_ENDS), smoke,
# a short bench line of both variants, then the driver's default bench line + reference arm + ncu launch list of the faster one.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( time KMAP_FOLD_RUN_ENDS=1 timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_fold.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_fold.log
tail -4 gpurun_out/pytest_gpu_fold.log
( time KMAP_FOLD_RUN_ENDS=1 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
SHORT="--steps 10 --warmup 3 --no-e2e --no-hamdist --no-piece2 --no-cpu --no-workflow"
( KMAP_FOLD_RUN_ENDS=0 timeout 200 python bench.py $SHORT ) > gpurun_out/bench_short_nofold.log 2> gpurun_out/bench_short_nofold.err
( KMAP_FOLD_RUN_ENDS=1 timeout 200 python bench.py $SHORT ) > gpurun_out/bench_short_fold.log 2> gpurun_out/bench_short_fold.err
FOLD=$(python - <<'P'
import json
def load(f):
    try:
        for l in open(f):
            if l.startswith('{'):
                return json.loads(l)
    except Exception:
        pass
    return None
a, b = load('gpurun_out/bench_short_nofold.log'), load('gpurun_out/bench_short_fold.log')
ok = bool(a and b and all(v is True or not isinstance(v, bool) for v in b['checks'].values())
          and b['checks']['table_checksums'] == a['checks']['table_checksums'] and b['ms_per_step'] < a['ms_per_step'])
import sys
for name, d in (('nofold', a), ('fold', b)):
    if d: print(name, round(d['ms_per_step'], 3), d['roofline']['phases_ms'], {k: v for k, v in d['checks'].items() if k != 'table_checksums'}, file=sys.stderr)
print(1 if ok else 0)
P
)
echo "KMAP_FOLD_RUN_ENDS=$FOLD for the full line" | tee gpurun_out/fold_choice.txt
export KMAP_FOLD_RUN_ENDS=$FOLD
( time timeout 400 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-hamdist --no-piece2 --no-workflow > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
