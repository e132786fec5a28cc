#!/bin/bash
# the driver's scaling launch at N ranks: default flags
mkdir -p gpurun_out
N=${1:-4}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
tail -c 600 gpurun_out/bench_${N}gpu.log; tail -4 gpurun_out/bench_${N}gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 ) > gpurun_out/bench_ref_${N}gpu.log 2>&1; tail -c 400 gpurun_out/bench_ref_${N}gpu.log
