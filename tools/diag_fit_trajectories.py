import os, sys
import numpy as np, torch
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT, "jmd-diversity-in-bayesian-optimization_b200")); sys.path.insert(0, ROOT)
import BO, objectives, data_gen
from oracle import gp as ogp
dev="cuda:0"
gen = objectives.objectives(); gen.device=dev
surf = gen.wildcatwells_batch([0,1,2,3], 0.2, 0.2, device=dev)
obj = surf[:, :100, :100].reshape(4, -1).contiguous()
obj_h = obj.cpu().numpy()
g = data_gen.diverse_data_generator(); g.device=dev
g.options["bounds"]=[(0,100),(0,100)]; g.options["N_samples"]=10000; g.gamma=1e-5; g.training_size=10
sets=[]
for s in range(4):
    g.options["seed"]=s; g.rng=None; g.generate(); sets.append(g.sample(95))
X0=np.stack(sets); sid=np.arange(4).astype(np.int32)
y0=obj_h[sid[:,None], (X0[:,:,1]*100+X0[:,:,0]).astype(int)].astype(np.float64)
r = BO.BOloop_batched(obj, sid, X0, y0, 40, theta=None, tol=0.1, device=dev)
torch.cuda.synchronize()
yall=r.yall.cpu().numpy(); th=r.thetas.cpu().numpy(); nit=r.nit.cpu().numpy(); picked=r.picked.cpu().numpy()
np.set_printoptions(precision=3, suppress=True, linewidth=200)
for t in range(4):
    print("trial",t,"nit",nit[t],"gap curve",(100-yall[t,:nit[t]+2]))
    for it in range(min(nit[t],12)):
        print("   it",it,"theta(noise,c,s,l)",th[t,it],"picked",picked[t,it]%100,picked[t,it]//100)
    # compare the fit at iteration 0 with scipy from default start
    raw, v_ref, res = ogp.fit_scipy(X0[t], y0[t], 3)
    print("   scipy fit at it0: theta", ogp.raw_to_theta(raw), "mll", v_ref)
