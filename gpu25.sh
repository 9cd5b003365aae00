ncu --set full --import-source on --clock-control none -k regex:perm_kernel -s 1 -c 1 -o gpurun_out/prof_r1_perm_v3 -f python gpu9.py > gpurun_out/b_ncu7.log 2>&1
tail -3 gpurun_out/b_ncu7.log
ls -la gpurun_out/prof_r1_perm_v3.ncu-rep
