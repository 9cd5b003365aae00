ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-perm > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fast_pair_kernel -s 3 -c 1 -o gpurun_out/prof_r1_fast3 python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > gpurun_out/b_ncu2.log 2>&1
