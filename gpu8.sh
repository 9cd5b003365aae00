python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.err; cat gpurun_out/bench_r1_d.json | cut -c1-1500; tail -3 gpurun_out/bench_r1_d.err
