ncu --set full --import-source on --clock-control none -k regex:perm_kernel -s 1 -c 1 -o gpurun_out/prof_r1_permall -f python gpu28.py > gpurun_out/b_ncu8.log 2>&1
tail -3 gpurun_out/b_ncu8.log
