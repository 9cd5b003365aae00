ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fast_pair_kernel -s 3 -c 1 -o gpurun_out/prof_r1_fast_final python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prep_x_dmma -s 3 -c 1 -o gpurun_out/prof_r1_dmma_final python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out
