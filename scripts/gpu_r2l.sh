#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "sort or radix or wide or stock_k or sorted or mixed" ) > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -8 gpurun_out/pytest_new.log
python scripts/bench_next.py 1e6 1e6 > gpurun_out/next_onesweep.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/next.json'))['sorted_path']
print({k: d[k] for k in d if k.startswith(('sorted_k16','phases'))})
PY
KMAP_SORT_THREE_KERNEL=1 python scripts/bench_next.py 1e6 1e6 > gpurun_out/next_three.log 2>&1; python - <<'PY'
import json
d=json.load(open('gpurun_out/next.json'))['sorted_path']
print('three-kernel', {k: d[k] for k in d if k.startswith(('sorted_k16','phases'))})
PY
