"""How many elements the exact cdf's chain walks serially (binade changes of the running sum), for a few weight families."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy.stats import poisson
import engine_helpers as eh
from bayesssm_b200 import _native as nat
ctx = nat.Context(0)
rng = np.random.default_rng(1)
for n in (1 << 18, 1 << 20):
    I = np.clip(70 + 25 * rng.standard_normal(n), 0, 500).round()
    fam = {"sir-like dpois(80 | I)": poisson.pmf(80, np.maximum(I, 1e-9)), "uniform": rng.random(n),
           "pf (many tiny)": np.exp(-0.5 * (rng.standard_normal(n) * 3) ** 2 * 4)}
    for name, w in fam.items():
        cdf, tot, ns = eh.cdf(ctx, w)
        print(f"n={n} {name}: serial elements {ns} ({ns / 1024:.1f} tiles of {n // 1024})", flush=True)
