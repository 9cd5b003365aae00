python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --no-perm 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-perm > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:fast_pair_kernel -s 3 -c 1 -o gpurun_out/prof_r1_fast_v3 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > gpurun_out/b_ncu5.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:prep_x_dmma -s 3 -c 1 -o gpurun_out/prof_r1_dmma_v3 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-perm > gpurun_out/b_ncu6.log 2>&1
ls -la gpurun_out/*.ncu-rep
