#!/bin/bash
# 2-GPU visit: folded run-end corrections under the one-exchange peer path (KMAP_FOLD_RUN_ENDS=2): the 2-rank count test + a short bench line
mkdir -p gpurun_out
export KMAP_FOLD_RUN_ENDS=2
( time timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "count_kmers" ) > gpurun_out/pytest_multi_fold.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_fold.log
tail -12 gpurun_out/pytest_multi_fold.log
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-hamdist --no-piece2 ) > gpurun_out/bench_peer_2gpu_fold.log 2> gpurun_out/bench_peer_2gpu_fold.err
python - <<P
import json
for l in open('gpurun_out/bench_peer_2gpu_fold.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['exchange'], d['scattered_merge']['ms_per_step'], d['roofline']['phases_ms'], d['checks'])
P
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_peer_2gpu_fold.err | tail -5
