import sys,os,time,subprocess
sys.path.insert(0,'.')
import bench, eqtlbma_b200
lib=eqtlbma_b200.load_library()
ds,_=bench.make_shard(0,1,lib,None)
sub=bench.subset_dataset(ds,500)
tmp='/tmp/cli500'; sub.write_files(tmp)
exe='eqtlbma_b200/eqtlbma_bf'
args=sub.ref_args(tmp,tmp+'/o')+bench.REF_FLAGS+['--thread','16','-v','1']
for i in range(3):
    t=time.perf_counter(); r=subprocess.run([exe]+args,capture_output=True,text=True); print("run",i,round(time.perf_counter()-t,3),r.returncode)
for thr in ('1','4'):
    a=[x for x in args]; a[a.index('--thread')+1]=thr
    t=time.perf_counter(); r=subprocess.run([exe]+a,capture_output=True,text=True); print('threads',thr,round(time.perf_counter()-t,3))
a=[x for x in args if x not in('--outss',)]
t=time.perf_counter(); r=subprocess.run([exe]+a,capture_output=True,text=True); print('no outss',round(time.perf_counter()-t,3))
a=[x for x in args]; a[a.index('-v')+1]='2'
r=subprocess.run([exe]+a,capture_output=True,text=True); print([l for l in r.stdout.splitlines() if l.startswith('phases')])
