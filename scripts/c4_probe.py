"""Configuration C4 (chain-binomial SIR, N = 2^18, general kernels): one APF / RMPF / BPF run each, for launch lists."""
import sys, time
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import engine_helpers as eh
from bayesssm_b200 import _native as nat

ctx = nat.Context(0)
rng = np.random.default_rng(4)
S, I, ys = 430, 70, []
T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for _ in range(T):
    ni = rng.binomial(S, 1 - np.exp(-0.5 * I / 500)); nr = rng.binomial(I, 1 - np.exp(-0.2))
    S, I = S - ni, I + ni - nr
    ys.append(rng.poisson(max(I, 0)))
y = np.array(ys, dtype=float)
for alg, name in ((1, "apf"), (2, "rmpf"), (0, "bpf")):
    for prec, pn in ((nat.F32, "f32"), (nat.F64, "f64")):
        best = 1e9
        for rep in range(3):
            r = eh.filter_run(ctx, 3, alg, 2, 0, 1 << 18, y, [0.5, 0.2, 500.0, 70.0], seed=7, precision=prec)
            best = min(best, float(r["kernel_ms"]))
        print(f"C4 {name} {pn}: {best:.3f} ms, {(1 << 18) * T / best / 1e6:.2f} G particle-timesteps/s, loglike {r['loglike'][0]:.3f}", flush=True)
