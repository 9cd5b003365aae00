#!/bin/bash
# ncu --set full capture of one kernel of the c2 bench step, raw-page CSV only; $1 = kernel regex, $2 = output stem
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:$1 -s 4 -c 1 -o gpurun_out/$2 \
    python bench.py --no-cpu --no-perm --no-e2e --steps 2 --warmup 3 > gpurun_out/$2.log 2>&1
ncu -i gpurun_out/$2.ncu-rep --page raw --csv > gpurun_out/$2_raw.csv 2>/dev/null
rm -f gpurun_out/$2.ncu-rep
