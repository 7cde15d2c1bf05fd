"""Throughput of the BASELINE.json configurations that are not the bench.py headline (device time of the engine
launch, CUDA events, best of 3): C2 with the three resamplers, C3 (256 LG filters vs Kalman), C4 (SIR, APF + RMPF).
Prints one JSON line per configuration.  python scripts/bench_configs.py > profiles/r1_bench_configs.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import engine_helpers as eh  # noqa: E402
import oracle  # noqa: E402  (Kalman value of C3: the checker, not the thing measured)
from bayesssm_b200 import _native as nat  # noqa: E402
from bench import simulate_y  # noqa: E402

ctx = nat.Context(0)


def best(fn, reps=3):
    out, ms = None, None
    for _ in range(reps):
        r = fn()
        if ms is None or r["kernel_ms"] < ms:
            out, ms = r, r["kernel_ms"]
    return out, ms


def emit(name, C, N, T, ms, extra):
    print(json.dumps({"config": name, "filters": C, "N": N, "T": T, "ms": ms,
                      "particle_timesteps_per_s": C * N * T / (ms * 1e-3), **extra}), flush=True)


# C2: bootstrap filter, nonlinear AR, T = 1000, N = 2^20, SISAR threshold 0.5 N, three resamplers
y = simulate_y(1000)
for rfn, name in ((0, "stratified"), (1, "systematic"), (2, "multinomial")):
    r, ms = best(lambda: eh.filter_run(ctx, 0, 0, 2, rfn, 1 << 20, y, [0.8, 1.0, 0.5], threshold=0.5 * (1 << 20), seed=1405,
                                       precision=nat.F32))
    emit(f"C2 bootstrap_filter nonlinear-AR SISAR {name}", 1, 1 << 20, 1000, ms,
         {"loglike": float(r["loglike"][0]), "n_resampled": int(r["n_resampled"][0]),
          "engine": "persistent" if rfn < 2 else "general (inverse-cdf multinomial)"})

# C3: linear-Gaussian, T = 500, N = 2^16, 256 batched filters, SISR, against the exact Kalman log-likelihood
rng = np.random.default_rng(3)
x, ys = rng.standard_normal(), []
for _ in range(500):
    x = 0.8 * x + rng.standard_normal()
    ys.append(x + rng.standard_normal())
y3 = np.array(ys)
r, ms = best(lambda: eh.filter_run(ctx, 1, 0, 1, 0, 1 << 16, y3, [0.8, 1.0, 1.0], seed=1405, num_filters=256, precision=nat.F32))
lls = r["loglike"]
est = float(np.log(np.mean(np.exp(lls - lls.max()))) + lls.max())
exact = float(oracle.kalman_loglik(y3, 0.8, 1.0, 1.0))
se = float(lls.std(ddof=1) / np.sqrt(len(lls)))
emit("C3 linear-Gaussian 256 filters SISR stratified", 256, 1 << 16, 500, ms,
     {"logmeanexp_loglike": est, "kalman_loglike": exact, "mc_standard_error": se, "abs_diff_in_se": abs(est - exact) / se})

# C4: stochastic SIR (chain-binomial), Poisson observations, T = 100, N = 2^18, APF and RMPF (general kernels)
rng = np.random.default_rng(4)
S, I, ys = 430, 70, []
for _ in range(100):
    ni = rng.binomial(S, 1 - np.exp(-0.5 * I / 500))
    nr = rng.binomial(I, 1 - np.exp(-0.2))
    S, I = S - ni, I + ni - nr
    ys.append(rng.poisson(max(I, 0)))
y4 = np.array(ys, dtype=float)
for alg, name in ((1, "auxiliary_filter"), (2, "resample_move_filter"), (0, "bootstrap_filter")):
    for prec, pn in ((nat.F64, "f64"), (nat.F32, "f32")):
        r, ms = best(lambda: eh.filter_run(ctx, 3, alg, 2, 0, 1 << 18, y4, [0.5, 0.2, 500.0, 70.0], seed=7, precision=prec))
        emit(f"C4 SIR {name} {pn}", 1, 1 << 18, 100, ms, {"loglike": float(r["loglike"][0]), "n_resampled": int(r["n_resampled"][0])})
