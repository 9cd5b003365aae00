set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
for k in fast_pair_kernel prep_x_dmma prep_y_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o /tmp/prof_$k -f python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > /dev/null 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > gpurun_out/r1_raw_$k.csv 2>/dev/null
done
ncu --set full --import-source on --clock-control none -k regex:perm_kernel -s 1 -c 1 -o /tmp/prof_perm -f python gpu9.py > /dev/null 2>&1
ncu -i /tmp/prof_perm.ncu-rep --page raw --csv > gpurun_out/r1_raw_perm_kernel_gensin.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:perm_kernel -s 1 -c 1 -o /tmp/prof_permall -f python gpu28.py > /dev/null 2>&1
ncu -i /tmp/prof_permall.ncu-rep --page raw --csv > gpurun_out/r1_raw_perm_kernel_all.csv 2>/dev/null
ls -la gpurun_out/r1_*
