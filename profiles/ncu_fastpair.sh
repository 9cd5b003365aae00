#!/bin/bash
# ncu --set full capture (with source) of the K2+K3 kernel of the c2 bench step; $1 = kernel regex, $2 = output stem
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s 4 -c 1 -o gpurun_out/$2 \
    python bench.py --no-cpu --no-perm --no-e2e --steps 2 --warmup 3 > gpurun_out/$2.log 2>&1
ncu -i gpurun_out/$2.ncu-rep --page raw --csv > gpurun_out/$2_raw.csv 2>/dev/null
ncu -i gpurun_out/$2.ncu-rep --page source --csv > gpurun_out/$2_source.csv 2>/dev/null
rm -f gpurun_out/$2.ncu-rep; ls -la gpurun_out/$2*
