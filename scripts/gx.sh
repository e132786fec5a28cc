mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "count or partition or streamed or find_motif_small" 2>&1 | tail -2
python scripts/phases.py 1e8 check 2>&1 | tee gpurun_out/phases.log
