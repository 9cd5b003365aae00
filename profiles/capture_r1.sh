#!/bin/bash
# Round-1 evidence run (under gpurun, one B200): GPU tests, smoke, both bench arms, launch list of the bench command,
# ncu --set full capture of the dominant kernel (raw + source pages as CSV), the c3/c4 slices.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r1_tests_gpu.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r1_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1_launches_bench_c2.log 2>&1
bash profiles/ncu_fastpair.sh fast_pair_warp_kernel r1_fast_pair_warp_kernel > /dev/null 2>&1
python profiles/gtex_slice.py > gpurun_out/r1_gtex_slice.log 2>&1
grep -E "passed|failed" gpurun_out/r1_tests_gpu.log; tail -n 1 gpurun_out/r1_smoke.log; cat gpurun_out/r1_bench_n1.json gpurun_out/r1_bench_reference.json; cat gpurun_out/r1_gtex_slice.log
