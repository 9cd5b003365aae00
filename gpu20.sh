python -m pytest tests -m gpu -q -x 2>&1 | tail -5
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-perm --verbose 2>gpurun_out/verbose_$1.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2))"; tail -5 gpurun_out/verbose_$1.err; }
run default
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-perm > /dev/null 2>&1
grep -c . gpurun_out/launches_r1g.csv
