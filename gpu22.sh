python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { python bench.py --steps 10 --warmup 3 --no-cpu $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2), d.get('perm',{}).get('permuted_pairs_per_s'))"; }
run base
cp variants/lb4.so eqtlbma_b200/libeqtlbma_b200.so
run lb4 --no-perm
cp variants/base.so eqtlbma_b200/libeqtlbma_b200.so
run base2 --no-perm
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-perm > /dev/null 2>&1
