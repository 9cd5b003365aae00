#!/bin/bash
# Round-1 evidence run (under gpurun, one B200): launch list of the bench command, the c3/c4 slices, the bench line.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r1_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1_launches_bench_c2.log 2>&1
python profiles/gtex_slice.py > gpurun_out/r1_gtex_slice.log 2>&1
python profiles/perm_slice_gensin.py > gpurun_out/r1_perm_slice_gensin.log 2>&1
tail -8 gpurun_out/r1_gtex_slice.log gpurun_out/r1_perm_slice_gensin.log
