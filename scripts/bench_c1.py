"""BASELINE config C1: README nonlinear-AR model, T = 20, pmmh with 4 chains, m = 1000, default tune_control
(pilot 2000 iterations at N = 100, 100 replicate filters) -- the regime the reference itself runs in.
Times the GPU engine (both precisions) and one chain of the CPU oracle."""
import sys, time, warnings
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import bayesssm_b200 as b
import oracle

rng = np.random.default_rng(1405)
x, ys = rng.standard_normal(), []
for _ in range(20):
    x = 0.8 * x + np.sin(x) + rng.standard_normal()
    ys.append(x + 0.5 * rng.standard_normal())
y = np.array(ys)
m = b.models.nonlinear_ar()
pri = {"phi": b.priors.uniform(0, 1), "sigma_x": b.priors.exponential(1), "sigma_y": b.priors.exponential(1)}
for chains in (4, 64, 1024):
    init = [{"phi": 0.8, "sigma_x": 1.0, "sigma_y": 0.5}] * chains
    for prec in ("f64", "f32"):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            out = b.pmmh(b.bootstrap_filter, y, 1000, m.init_fn, m.transition_fn, m.log_likelihood_fn, pri, init, burn_in=200,
                         num_chains=chains, param_transform={"phi": "logit", "sigma_x": "log", "sigma_y": "log"},
                         seed=1405, print_result=False, precision=prec)
            dt = time.perf_counter() - t0
        tc = out["theta_chain"]
        print(f"chains={chains:5d} {prec}: wall {dt:7.2f} s  pilot {out['timing_ms']['pilot']:9.1f} ms  main {out['timing_ms']['main']:9.1f} ms  "
              f"target_n {int(out['target_n'].min())}-{int(out['target_n'].max())}  phi {tc['phi'].mean():.3f}  sigma_x {tc['sigma_x'].mean():.3f}  "
              f"sigma_y {tc['sigma_y'].mean():.3f}  acc {out['acceptance_rate'].mean():.3f}", flush=True)
t0 = time.perf_counter()
r = oracle.pmmh_chain(0, 0, y, [0.8, 1.0, 0.5], [3, 2, 2], [0, 1, 1], [1, 0, 0], [2, 1, 1], [0.5] * 3, 100, 2000, 100, 1000, 0, 1405)
print(f"CPU oracle, ONE chain: {time.perf_counter() - t0:.2f} s  target_n {r['target_n']}  phi {r['theta_chain'][200:, 0].mean():.3f}")
