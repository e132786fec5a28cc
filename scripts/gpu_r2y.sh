#!/bin/bash
# N-GPU visit: the exchange step alone (peer memory vs NCCL), then the bench line with 2 / 1 key ranges
mkdir -p gpurun_out
N=${1:-8}
( KMAP_PEER_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/bench_exchange.py ) > gpurun_out/exchange_${N}gpu.log 2> gpurun_out/exchange_${N}gpu.err
grep "ranks" gpurun_out/exchange_${N}gpu.log; grep "peer trace] rank 0" gpurun_out/exchange_${N}gpu.err | tail -3
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/bench_exchange.py ) > gpurun_out/exchange_${N}gpu_notrace.log 2>&1
grep "ranks" gpurun_out/exchange_${N}gpu_notrace.log
for C in 2 1; do
( KMAP_MERGE_CHUNKS=$C timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-hamdist --no-piece2 ) > gpurun_out/bench_peer_${N}gpu_c$C.log 2> gpurun_out/bench_peer_${N}gpu_c$C.err
python - <<P
import json
for l in open('gpurun_out/bench_peer_${N}gpu_c$C.log'):
    if l.startswith('{'):
        d=json.loads(l); print('chunks $C', d['value'], d['ms_per_step'], d['exchange'], d['scattered_merge']['ms_per_step'], d['roofline']['phases_ms'], d['checks'])
P
tail -3 gpurun_out/bench_peer_${N}gpu_c$C.err
done
