#!/bin/bash
# 8-GPU visit: a traced short run (where the exchange sits in the step), then the driver's default launch
mkdir -p gpurun_out
N=${1:-8}
( KMAP_MERGE_TRACE=1 KMAP_PEER_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-hamdist --no-piece2 ) > gpurun_out/bench_peer_${N}gpu_trace.log 2> gpurun_out/bench_peer_${N}gpu_trace.err
grep "merge trace\|peer trace] rank 0" gpurun_out/bench_peer_${N}gpu_trace.err | sed -n 17,26p
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
python - <<P
import json
for l in open('gpurun_out/bench_${N}gpu.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['exchange'], d['scattered_merge']['ms_per_step'], d['roofline']['phases_ms'], d['e2e']['ms_per_step'], d['checks'])
P
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_${N}gpu.err | tail -5
