"""C4 in the parity precision only (for launch lists): APF f64, N = 2^18."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import engine_helpers as eh
from bayesssm_b200 import _native as nat
ctx = nat.Context(0)
rng = np.random.default_rng(4)
y = rng.poisson(80, 10).astype(float)
r = eh.filter_run(ctx, 3, 1, 2, 0, 1 << 18, y, [0.5, 0.2, 500.0, 70.0], seed=7, precision=nat.F64)
print(float(r["kernel_ms"]), r["loglike"])
