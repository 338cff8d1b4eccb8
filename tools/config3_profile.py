import os, sys, time
import numpy as np, torch
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200")); sys.path.insert(0, ROOT)
from b200bo import engine, fit, ops
import objectives
dev=torch.device("cuda:0")
FAM=[(s,r) for s in (0.2,0.4,0.6,0.8) for r in (0.2,0.4,0.6,0.8)]
gen=objectives.objectives(); S=64
surf=gen.wildcatwells_batch([s//16 for s in range(S)],[FAM[s%16][0] for s in range(S)],[FAM[s%16][1] for s in range(S)],device=dev)
lat=ops.Lattice(100,100); pts=lat.points(dev); obj=surf[:,:100,:100].reshape(S,-1).contiguous()
rng=np.random.default_rng(99)
for tol in (0.1,-1.0):
    T=1000
    idx=torch.as_tensor(np.stack([rng.choice(10000,10,replace=False) for _ in range(T)]),device=dev)
    sid=torch.as_tensor((np.arange(T)%S).astype(np.int32),device=dev)
    X0=pts[idx]; y0=obj[sid.long()[:,None],idx].double()
    bo=engine.BatchedBO(pts,obj,sid,X0,y0,100,theta=None,tol=tol,fitter=fit.LBFGSFitter(T,dev),lattice=lat)
    bo.run(); bo.reset(X0,y0); bo.profile=[]
    torch.cuda.synchronize(); t0=time.time(); res=bo.run(); torch.cuda.synchronize(); dt=time.time()-t0
    agg={}
    for nm,n,e0,e1 in bo.profile: agg[nm]=agg.get(nm,0)+e0.elapsed_time(e1)
    act=[int((res.nit>it).sum()) for it in (0,10,20,30,50,80)]
    print("tol",tol,"total %.1f ms"%(dt*1e3), {k:round(v,1) for k,v in agg.items()}, "active trials at it 0,10,20,30,50,80:",act)
    fits=[(n,e0.elapsed_time(e1)) for nm,n,e0,e1 in bo.profile if nm=="fit"]
    print("   fit ms by n:", [(n,round(ms,2)) for n,ms in fits[::10]])
