#!/bin/bash
# 2-GPU visit: push form of the peer-memory exchange -- the 2-rank count test, the exchange alone (push / pull, traced), the bench line
mkdir -p gpurun_out
N=${1:-2}
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "count_kmers" ) > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
for M in push pull; do
( KMAP_PEER_MODE=$M KMAP_PEER_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/bench_exchange.py ) > gpurun_out/exchange_${N}gpu_$M.log 2> gpurun_out/exchange_${N}gpu_$M.err
grep "ranks" gpurun_out/exchange_${N}gpu_$M.log; grep "peer trace] rank 0" gpurun_out/exchange_${N}gpu_$M.err | tail -2
( KMAP_PEER_MODE=$M timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/bench_exchange.py ) > gpurun_out/exchange_${N}gpu_${M}_notrace.log 2>&1
grep "ranks" gpurun_out/exchange_${N}gpu_${M}_notrace.log
done
( KMAP_MERGE_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-hamdist --no-piece2 ) > gpurun_out/bench_peer_${N}gpu.log 2> gpurun_out/bench_peer_${N}gpu.err
python - <<P
import json
for l in open('gpurun_out/bench_peer_${N}gpu.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['exchange'], d['scattered_merge']['ms_per_step'], d['roofline']['phases_ms'], d['checks'])
P
grep -v "merge trace" gpurun_out/bench_peer_${N}gpu.err | grep -v "^\*\*\*\|OMP_NUM" | tail -5; grep "merge trace" gpurun_out/bench_peer_${N}gpu.err | sed -n 61,70p
