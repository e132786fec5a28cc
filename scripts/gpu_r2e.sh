#!/bin/bash
# N-GPU device-resident bench line for several settings of the SMs left to the exchange
mkdir -p gpurun_out
N=${1:-2}
for R in 0 8 16 32; do
( KMAP_MERGE_RESERVE_SMS=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e --no-hamdist ) > gpurun_out/bench_tmp.log 2> gpurun_out/bench_tmp.err
python - $R <<'PY'
import json,sys
l=[x for x in open('gpurun_out/bench_tmp.log') if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); print('reserve', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['phases_ms'].items()})
else: print('reserve', sys.argv[1], 'FAILED')
PY
done
