python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu --verbose > gpurun_out/bench_r1_c.json 2> gpurun_out/bench_r1_c.err; cat gpurun_out/bench_r1_c.json | cut -c1-1400; tail -3 gpurun_out/bench_r1_c.err
python bench.py --steps 5 --warmup 3 --no-cpu --no-perm --verbose 2>&1 | grep -E "e2e host|e2e" | cut -c1-400
