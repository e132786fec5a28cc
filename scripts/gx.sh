mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "hamdist or scan_motif" 2>&1 | tail -2
python scripts/hamdist_bench.py
