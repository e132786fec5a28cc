set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "count or partition or dedup or find_motif_small" > gpurun_out/pytest_count.log 2>&1; tail -3 gpurun_out/pytest_count.log
python scripts/phases.py 1e8 check > gpurun_out/phases.log 2>&1; cat gpurun_out/phases.log
