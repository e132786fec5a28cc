#!/bin/bash
# N-GPU bench line (device-resident step with the in-count all-reduce, e2e with sharded lists, distance matrix by row blocks)
mkdir -p gpurun_out
N=${1:-8}
nproc > gpurun_out/nproc_${N}gpu.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-piece2 --no-workflow ) > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
tail -c 1200 gpurun_out/bench_${N}gpu.log; tail -3 gpurun_out/bench_${N}gpu.err; cat gpurun_out/nproc_${N}gpu.txt
