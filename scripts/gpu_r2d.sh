#!/bin/bash
# 2-GPU visit: the 2-rank NCCL tests, then the 2-GPU bench line (device-resident + e2e)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider ) > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -30 gpurun_out/pytest_multi.log
N=${1:-2}
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu ) > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; tail -5 gpurun_out/bench_${N}gpu.err
python - $N <<'PY'
import json,sys
l=[x for x in open(f'gpurun_out/bench_{sys.argv[1]}gpu.log') if x.startswith('{"metric')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step']); print(d['roofline']['phases_ms']); print(d['e2e']); print(d['checks']); print(d.get('hamdist',{}).get('ms'))
PY
