ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:prep_x_dmma -s 18 -c 1 -o /tmp/prof_dmma -f python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > /dev/null 2>&1
ncu -i /tmp/prof_dmma.ncu-rep --page raw --csv > gpurun_out/r1_raw_prep_x_dmma.csv 2>/dev/null
grep -c . gpurun_out/r1_launches_bench_c2.csv
