import sys, numpy as np, torch
sys.path.insert(0,'.')
import bench
from vstrains_b200 import pe_inference
cfg,g,f,r = bench.make_workload("C2", 500000, 0)
ix = pe_inference.PEIndex([bytes(s) for s in g.seqs], cfg.k)
d_f = torch.from_numpy(f).cuda(); d_r = torch.from_numpy(r).cuda()
for _ in range(3):
    ix.reset(); ix.count_device(d_f.data_ptr(), f.size, d_r.data_ptr(), r.size)
ix.set_option("dbg_times", 1)
ix.reset(); ix.count_device(d_f.data_ptr(), f.size, d_r.data_ptr(), r.size)
nt = (r.size + 49151)//49152
buf = np.zeros((nt+4, 8), dtype=np.uint64)
ix.set_option("dbg_dump", buf.ctypes.data)
t = buf[:nt].astype(np.int64)
t0 = t[:,0].min()
t -= t0
print("tiles", nt, "kernel span us", (t[:,5].max())/1e3)
d = lambda a,b: (t[:,b]-t[:,a])/1e3
for name,a,b in (("tma wait",0,1),("M1+M2+agg",1,2),("lookback",2,3),("emit",3,4),("pack",4,5),("total",0,5)):
    x = d(a,b); print("%-10s mean %6.2f  p50 %6.2f  p90 %6.2f  p99 %6.2f  max %6.2f us"%(name,x.mean(),np.median(x),np.percentile(x,90),np.percentile(x,99),x.max()))
# concurrency: average number of tiles in flight
ev = np.concatenate([np.stack([t[:,0], np.ones(nt)],1), np.stack([t[:,5], -np.ones(nt)],1)])
ev = ev[np.argsort(ev[:,0])]
conc = np.cumsum(ev[:,1]); dt = np.diff(ev[:,0]); print("avg tiles in flight %.1f"%((conc[:-1]*dt).sum()/dt.sum()))
# start order vs tile id
print("start times of tiles 0,100,1000,last:", t[0,0]/1e3, t[100,0]/1e3, t[min(1000,nt-1),0]/1e3, t[nt-1,0]/1e3)
np.save("gpurun_out/dbg_times.npy", t)
