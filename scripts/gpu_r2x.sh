#!/bin/bash
# N-GPU visit for the peer-memory exchange: the 2-rank NCCL tests (count mode first), then the bench line at N ranks
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -30 gpurun_out/pytest_multi.log
( time KMAP_MERGE_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-hamdist --no-piece2 ) > gpurun_out/bench_peer_${N}gpu.log 2> gpurun_out/bench_peer_${N}gpu.err
tail -c 1500 gpurun_out/bench_peer_${N}gpu.log; grep -v "merge trace" gpurun_out/bench_peer_${N}gpu.err | tail -5; grep "merge trace" gpurun_out/bench_peer_${N}gpu.err | tail -8
