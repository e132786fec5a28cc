#!/bin/bash
# 2-GPU visit: coalesced 4-cell-unit exchange kernels -- the 2-rank count test, the exchange alone (traced), the bench line
mkdir -p gpurun_out
N=${1:-2}
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "count_kmers" ) > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
( KMAP_PEER_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/bench_exchange.py ) > gpurun_out/exchange_${N}gpu_push4.log 2> gpurun_out/exchange_${N}gpu_push4.err
grep "ranks" gpurun_out/exchange_${N}gpu_push4.log; grep "peer trace] rank 0" gpurun_out/exchange_${N}gpu_push4.err | tail -2
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-hamdist --no-piece2 ) > gpurun_out/bench_peer_${N}gpu.log 2> gpurun_out/bench_peer_${N}gpu.err
python - <<P
import json
for l in open('gpurun_out/bench_peer_${N}gpu.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['exchange'], d['scattered_merge']['ms_per_step'], d['roofline']['phases_ms'], d['checks'])
P
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_peer_${N}gpu.err | tail -5
