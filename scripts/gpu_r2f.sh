#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
( KMAP_MERGE_TRACE=1 KMAP_MERGE_RESERVE_SMS=${2:-16} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --no-e2e --no-hamdist ) > gpurun_out/bench_tmp.log 2> gpurun_out/bench_tmp.err
grep "merge trace" gpurun_out/bench_tmp.err | tail -20
