python - <<'PY'
import ctypes, eqtlbma_b200
lib = eqtlbma_b200.load_library()
out=(ctypes.c_double*5)()
print(lib.eqb_math_selftest(0, ctypes.c_int64(3000000), out), list(out))
PY
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2), d.get('perm',{}).get('permuted_pairs_per_s'))"
python gpu14.py 2>&1 | tail -5
