python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu --verbose 2>gpurun_out/verbose_e2e.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value']/1e6,2), d.get('perm',{}).get('permuted_pairs_per_s'))"
tail -4 gpurun_out/verbose_e2e.err
