python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; cat gpurun_out/bench_r1_final.json | cut -c1-2500
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_final_ref.json 2>&1; cat gpurun_out/bench_r1_final_ref.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fast_pair_kernel -s 3 -c 1 -o gpurun_out/prof_r1_fast_final python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:prep_x_dmma -s 3 -c 1 -o gpurun_out/prof_r1_dmma_final python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > gpurun_out/b_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perm_kernel -s 1 -c 1 -o gpurun_out/prof_r1_perm_final python gpu9.py > gpurun_out/b_ncu4.log 2>&1
