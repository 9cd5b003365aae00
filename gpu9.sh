python gpu9.py
ncu --set full --clock-control none --import-source on -k regex:perm_kernel -s 1 -c 1 -o gpurun_out/prof_r1_perm python gpu9.py > gpurun_out/b_ncu4.log 2>&1
