set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; tail -c 3000 gpurun_out/bench_r1_a.json; tail -5 gpurun_out/bench_r1_a.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1_ref.json 2>&1; cat gpurun_out/bench_r1_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu --genes 1000 > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 3 -c 1 -o gpurun_out/prof_r1_pair python bench.py --steps 1 --warmup 3 --no-cpu --no-perm --genes 1000 > gpurun_out/b_ncu2.log 2>&1
tail -3 gpurun_out/b_ncu2.log
ls -la gpurun_out
