import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import engine_helpers as eh
from bayesssm_b200 import _native as nat
ctx = nat.Context(0)
rng = np.random.default_rng(4)
S, I, ys = 430, 70, []
for _ in range(100):
    ni = rng.binomial(S, 1 - np.exp(-0.5 * I / 500)); nr = rng.binomial(I, 1 - np.exp(-0.2))
    S, I = S - ni, I + ni - nr; ys.append(rng.poisson(max(I, 0)))
y = np.array(ys, dtype=float)
for alg in (1, 0):
    for N in (1 << 12, 1 << 15, 1 << 18):
        res = {}
        for prec, pn in ((nat.F64, "f64"), (nat.F32, "f32")):
            lls = [eh.filter_run(ctx, 3, alg, 2, 0, N, y, [0.5, 0.2, 500.0, 70.0], seed=s, precision=prec)["loglike"][0] for s in range(12)]
            res[pn] = (np.mean(lls), np.std(lls, ddof=1))
        print("alg", alg, "N", N, {k: (round(v[0], 3), round(v[1], 3)) for k, v in res.items()}, flush=True)
