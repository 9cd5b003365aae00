run() { python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2), d.get('perm',{}).get('permuted_pairs_per_s'))"; python gpu14.py 2>&1 | tail -4; }
run fast
cp variants/libm.so eqtlbma_b200/libeqtlbma_b200.so
run libm
cp variants/base.so eqtlbma_b200/libeqtlbma_b200.so
