python -m pytest tests/test_cli_dropin.py -q 2>&1 | tail -4
# CLI end-to-end timing on a c2-shaped sample (500 genes): ours vs the reference binary
python - <<'PY'
import sys, time, subprocess, os; sys.path.insert(0,'.')
import bench
from eqtlbma_b200.synth import make_dataset
wl = dict(bench.WORKLOAD, n_genes=500)
ds = make_dataset(**wl); d='/tmp/cli_c2'; ds.write_files(d)
flags=["--analys","join","--bfs","sin","--outss","--outw","-v","0"]
for thr in (1,16):
    t=time.time(); r=subprocess.run(["eqtlbma_b200/eqtlbma_bf"]+ds.ref_args(d,d+"/ours%d"%thr)+flags+["--thread",str(thr)],capture_output=True,text=True); print('ours thread',thr,round(time.time()-t,2),'s rc',r.returncode, r.stderr[-200:])
sub = bench.subset_dataset(ds, 60); d2='/tmp/cli_c2_ref'; sub.write_files(d2)
t=time.time(); r=subprocess.run(["oracle/_ref/eqtlbma_bf_ref"]+sub.ref_args(d2,d2+"/ref")+flags,capture_output=True,text=True); print('reference (60 genes only)',round(time.time()-t,2),'s rc',r.returncode)
PY
