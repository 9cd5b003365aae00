python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0,'.')
import bench, eqtlbma_b200
from eqtlbma_b200._capi import Engine as E
from eqtlbma_b200.synth import make_dataset
ds = bench.pinned_copy(make_dataset(**bench.WORKLOAD))
kw = dict(analysis="join", bfs="sin")
eng = eqtlbma_b200.Engine(ds, **kw)
out = eng.alloc_results(raw=True, pinned=True)
for it in range(3):
    E.timing = {}
    t0=time.perf_counter()
    e = eqtlbma_b200.Engine(ds, **kw)
    t1=time.perf_counter()
    e.run(raw=True, out=out)
    t2=time.perf_counter()
    e.close()
    t3=time.perf_counter()
    print('init %.1f ms run %.1f ms close %.1f ms'%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3), {k: round(v*1e3,2) for k,v in E.timing.items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
