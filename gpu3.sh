python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; cat gpurun_out/bench_r1_b.json; tail -5 gpurun_out/bench_r1_b.err
