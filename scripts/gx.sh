mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 2 --warmup 3 --no-cpu --extras > gpurun_out/bench_x.log 2> gpurun_out/bench_x.err; tail -3 gpurun_out/bench_x.err
