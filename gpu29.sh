python gpu28.py
python - <<'PY'
import ctypes, eqtlbma_b200
lib = eqtlbma_b200.load_library()
out=(ctypes.c_double*5)()
print(lib.eqb_math_selftest(0, ctypes.c_int64(3000000), out), list(out))
PY
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python gpu14.py 2>&1 | tail -3
