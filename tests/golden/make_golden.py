"""Generates tests/golden/*.npz.

The reference (an R package) cannot be executed in this image, so these vectors do NOT come from it: they are
outputs of the CPU oracle (oracle/pf_oracle.c) on fixed inputs, committed so that (a) the oracle itself cannot drift
silently and (b) the CUDA engine is checked against stored numbers as well as against the live oracle.  Inputs that
have a hand-derivable answer (tie rule, degenerate weights, the structural cases of tests/testthat/test-resampling.R)
are stored with that answer.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402


def main():
    rng = np.random.default_rng(1405)
    # ---- resampling: the reference's own test weights + random cases, fixed uniforms ----
    cases = {}
    w_ref = np.array([0.1, 0.5, 0.1, 0.15, 0.15])          # tests/testthat/test-resampling.R:49
    w_prop = np.array([0.1, 0.2, 0.3, 0.2, 0.2])           # :31
    for name, w in (("ref", w_ref), ("prop", w_prop), ("degenerate", np.array([0, 0, 1.0, 0, 0])),
                    ("uniform4", np.full(4, 0.25)), ("random1000", rng.random(1000)),
                    ("pf4096", np.exp(-0.5 * (3 * rng.standard_normal(4096)) ** 2))):
        n = len(w)
        u = rng.random(n)
        if name == "uniform4":
            u = np.zeros(n)                                 # tie rule: pos == c[j] selects j => 1 1 2 3
        cases[f"{name}_w"] = w
        cases[f"{name}_u"] = u
        for kind in ("stratified", "systematic", "multinomial"):
            cases[f"{name}_{kind}"] = oracle.resample(kind, w, u)
        c, tot = oracle.cdf(w)
        cases[f"{name}_cdf"] = c
        cases[f"{name}_total"] = np.array([tot])
    np.savez(os.path.join(HERE, "resampling.npz"), **cases)

    # ---- filters: config C1 shape (README model, T = 20, N = 1000), injected noise ----
    T, N = 20, 1000
    x, ys = rng.standard_normal(), []
    for _ in range(T):
        x = 0.8 * x + np.sin(x) + rng.standard_normal()
        ys.append(x + 0.5 * rng.standard_normal())
    y = np.array(ys)
    out = {"y": y, "theta": np.array([0.8, 1.0, 0.5])}
    noise = oracle.make_noise(0, N, T, T, rng)
    for k, v in noise.items():
        out[f"noise_{k}"] = v
    for alg, aname in ((0, "bpf"), (1, "apf"), (2, "rmpf")):
        r = oracle.particle_filter(0, alg, 2, 0, N, y, [0.8, 1.0, 0.5], noise=noise, want_ancestors=True)
        out[f"{aname}_loglike"] = np.array([r["loglike"]])
        out[f"{aname}_loglike_history"] = r["loglike_history"]
        out[f"{aname}_ess"] = r["ess"]
        out[f"{aname}_state_est"] = r["state_est"]
        out[f"{aname}_ancestors"] = r["ancestors_history"]
    # Philox mode (no injected buffers): pins the counter-based generator on both sides
    r = oracle.particle_filter(0, 0, 2, 0, N, y, [0.8, 1.0, 0.5], seed=1405, run_id=7, stream=3)
    out["philox_loglike"] = np.array([r["loglike"]])
    out["philox_state_est"] = r["state_est"]
    np.savez_compressed(os.path.join(HERE, "filter_c1.npz"), **out)

    # ---- PMMH: one short chain of config C1 (README priors), Philox ----
    kw = dict(prior_kind=[3, 2, 2], prior_a=[0.0, 1.0, 1.0], prior_b=[1.0, 0.0, 0.0], transform=[2, 1, 1],
              pilot_proposal_sd=[0.1, 0.15, 0.2], pilot_n=64, pilot_m=30, pilot_reps=6, m=40, seed=99)
    r = oracle.pmmh_chain(0, 0, y[:12], [0.8, 1.0, 0.5], chain_id=4, **kw)
    np.savez(os.path.join(HERE, "pmmh_c1.npz"), y=y[:12], theta_chain=r["theta_chain"], loglike_chain=r["loglike_chain"],
             pilot_theta_mean=r["pilot_theta_mean"], pilot_theta_cov=r["pilot_theta_cov"],
             proposal_chol=r["proposal_chol"], target_n=np.array([r["target_n"]]), n_accept=np.array([r["n_accept"]]))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
