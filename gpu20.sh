python -m pytest tests -m gpu -q -x 2>&1 | tail -5
run() { python bench.py --steps 10 --warmup 3 --no-cpu --no-perm 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2))"; }
run default
EQB_FAST_T=16 run T16
EQB_FAST_T=64 EQB_FAST_SMEM_KB=110 run T64
EQB_FAST_T=32 run T32
