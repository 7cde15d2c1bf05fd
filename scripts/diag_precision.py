import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, engine_helpers as eh
from bayesssm_b200 import _native as nat
from test_filter_gpu import THETA, sim_y
ctx=nat.Context(0)
rng=np.random.default_rng(9); y=sim_y(0,20,rng)
for N in (1<<16, 1<<20):
    res={}
    for name,prec,eng in (("pers32",nat.F32,nat.ENGINE_PERSISTENT),("gen32",nat.F32,nat.ENGINE_GENERAL),("pers64",nat.F64,nat.ENGINE_PERSISTENT),("gen64",nat.F64,nat.ENGINE_GENERAL)):
        r=eh.filter_run(ctx,0,0,2,0,N,y,THETA[0],seed=5,precision=prec,engine=eng)
        res[name]=r
        print(N,name,r["loglike"][0],r["n_resampled"][0])
    for name in ("pers32","gen32","pers64"):
        d=res[name]["loglike_history"][0]-res["gen64"]["loglike_history"][0]
        print(name,"cum diff vs gen64:",np.round(d,5)[[0,1,2,3,5,9,14,19]], "ess diff", np.abs(res[name]["ess"][0]/res["gen64"]["ess"][0]-1).max())
