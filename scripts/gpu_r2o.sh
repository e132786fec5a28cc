#!/bin/bash
# N-GPU sweep of the exchange knobs: CTAs the collective may use, key ranges the level-14 table is merged in
mkdir -p gpurun_out
N=${1:-8}
: > gpurun_out/merge_sweep_${N}gpu.txt
for cfg in "16 4" "16 1" "32 1" "32 2" "64 1"; do
set -- $cfg
( KMAP_COMM_CTAS=$1 KMAP_MERGE_CHUNKS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --no-hamdist --no-piece2 --no-workflow ) > gpurun_out/bench_tmp.log 2> gpurun_out/bench_tmp.err
python - $1 $2 <<'PY' | tee -a gpurun_out/merge_sweep_${N}gpu.txt
import json,sys
l=[x for x in open('gpurun_out/bench_tmp.log') if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); s=d.get('scattered_merge') or {}
    print('ctas', sys.argv[1], 'chunks', sys.argv[2], 'allreduce ms/step', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['phases_ms'].items()}, 'scattered ms/step', round(s.get('ms_per_step',0),2), s.get('owned_ranges_equal_allreduced_tables'))
else: print('ctas', sys.argv[1], 'chunks', sys.argv[2], 'FAILED')
PY
done
