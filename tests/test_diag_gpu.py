"""GPU parity of the device MCMC diagnostics (bssm_mcmc_diagnostics, through the C ABI) against the numpy
restatement of R/ESS.R:30-104 and R/rhat.R:27-67 (oracle/mcmc_diag.py), and their use by pmmh() (R/pmmh.R:570-594).
Tolerance: 1e-10 relative (same sums in the same order; only the contraction of a*b+c into an FMA differs)."""
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mcmc_diag as od  # noqa: E402

import bayesssm_b200 as b  # noqa: E402
from bayesssm_b200.diagnostics import device_diagnostics  # noqa: E402

pytestmark = pytest.mark.gpu


def ar1(rng, m, k, rho):
    x = np.zeros((m, k))
    x[0] = rng.standard_normal(k)
    for t in range(1, m):
        x[t] = rho * x[t - 1] + rng.standard_normal(k)
    return x


@pytest.mark.parametrize("k,m,p,burn", [(2, 2, 1, 0), (3, 7, 2, 0), (4, 257, 3, 56), (16, 101, 2, 1), (8, 2000, 3, 500),
                                        (300, 64, 1, 0)])
def test_device_diagnostics_match_oracle(engine, k, m, p, burn):
    rng = np.random.default_rng(100 * k + m)
    draws = np.stack([ar1(rng, m, p, rho) * (1 + c % 5) + c % 3 for c, rho in zip(range(k), np.linspace(0.0, 0.95, k))], axis=0)
    draws[..., -1] *= 1e-3
    r = device_diagnostics(draws, burn, ctx=engine)
    for j in range(p):
        mat = draws[:, burn:, j].T
        if m - burn >= 4:
            np.testing.assert_allclose(r["rhat"][j], od.rhat_matrix(mat), rtol=1e-10)
        np.testing.assert_allclose(r["ess"][j], od.ess_matrix(mat), rtol=1e-10)
    assert (r["flags"] == 0).all()
    again = device_diagnostics(draws, burn, ctx=engine)
    np.testing.assert_array_equal(again["ess"], r["ess"])           # no atomics, fixed summation order


def test_public_ess_rhat(engine):
    import pandas as pd
    rng = np.random.default_rng(1405)
    iid = rng.standard_normal((1000, 3))
    assert abs(b.ess(iid, ctx=engine) - 3000) < 0.05 * 3000          # tests/testthat/test-ESS.R:1-5
    np.testing.assert_allclose(b.ess(iid, ctx=engine), od.ess_matrix(iid), rtol=1e-10)
    ar = ar1(rng, 1000, 3, 0.9)
    assert b.ess(ar, ctx=engine) < 3000                              # test-ESS.R:7-22
    assert b.rhat(iid, ctx=engine) < 1.01                            # test-rhat.R:1-5
    drift = np.concatenate([rng.standard_normal(50), rng.standard_normal(50) + 10])[:, None]
    assert b.rhat(drift, ctx=engine) > 2                             # test-rhat.R:18-27 (one chain)
    assert b.rhat(iid[:999], ctx=engine) < 1.01                      # test-rhat.R:64-69 (odd length)
    df = pd.DataFrame({"chain": np.repeat([1, 2, 3], 1000), "param1": iid.T.ravel(), "param2": ar.T.ravel()})
    e, r = b.ess(df, ctx=engine), b.rhat(df, ctx=engine)             # test-ESS.R:24-33, test-rhat.R:7-15
    np.testing.assert_allclose([e["param1"], e["param2"]], [od.ess_matrix(iid), od.ess_matrix(ar)], rtol=1e-10)
    np.testing.assert_allclose([r["param1"], r["param2"]], [od.rhat_matrix(iid), od.rhat_matrix(ar)], rtol=1e-10)
    with pytest.warns(UserWarning, match="One or more chains have zero variance"):
        assert np.isnan(b.ess(np.ones((3, 3)), ctx=engine))          # test-ESS.R:53-56
    with pytest.warns(UserWarning, match="One or more chains have zero variance"):
        assert np.isnan(b.rhat(np.ones((4, 4)), ctx=engine))         # test-rhat.R:42-45
    near = np.tile(np.array([0.0, 1.0] * 50)[:, None], (1, 2))
    assert b.rhat(near, ctx=engine) == 1.0                           # R/rhat.R:63-65


def test_pmmh_reports_device_diagnostics(engine):
    rng = np.random.default_rng(1405)
    x, ys = rng.standard_normal(), []
    for _ in range(12):
        x = 0.8 * x + np.sin(x) + rng.standard_normal()
        ys.append(x + 0.5 * rng.standard_normal())
    mdl = b.models.nonlinear_ar()
    pri = {"phi": b.priors.uniform(0, 1), "sigma_x": b.priors.exponential(1), "sigma_y": b.priors.exponential(1)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = b.pmmh(b.bootstrap_filter, np.array(ys), m=120, init_fn=mdl.init_fn, transition_fn=mdl.transition_fn,
                     log_likelihood_fn=mdl.log_likelihood_fn, log_priors=pri,
                     pilot_init_params=[{"phi": .8, "sigma_x": 1., "sigma_y": .5}] * 3, burn_in=20, num_chains=3,
                     tune_control=dict(b.default_tune_control(), pilot_m=40, pilot_n=50, pilot_reps=5),
                     seed=5, ctx=engine, print_result=False)
    tc = res["theta_chain"]
    for name in ("phi", "sigma_x", "sigma_y"):
        mat = np.stack([tc.loc[tc["chain"] == str(c + 1), name].to_numpy() for c in range(3)], axis=1)
        assert mat.shape == (100, 3)
        np.testing.assert_allclose(res["diagnostics"]["ess"][name], od.ess_matrix(mat), rtol=1e-10)
        np.testing.assert_allclose(res["diagnostics"]["rhat"][name], od.rhat_matrix(mat), rtol=1e-10)


def test_diagnostics_at_the_chain_count_of_config_c5(engine):
    """1024 chains x 1000 draws x 3 parameters (BASELINE.json configs[4]): one call, checked for one parameter."""
    rng = np.random.default_rng(7)
    k, m, p = 1024, 1000, 3
    draws = np.empty((k, m, p))
    draws[:, 0] = rng.standard_normal((k, p))
    eps = rng.standard_normal((k, m, p))
    for t in range(1, m):
        draws[:, t] = 0.7 * draws[:, t - 1] + eps[:, t]
    r = device_diagnostics(draws, 0, ctx=engine)
    mat = draws[:, :, 1].T
    np.testing.assert_allclose(r["ess"][1], od.ess_matrix(mat), rtol=1e-9)
    np.testing.assert_allclose(r["rhat"][1], od.rhat_matrix(mat), rtol=1e-10)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "diag_c5_timing.txt"), "w") as f:
            f.write(f"bssm_mcmc_diagnostics k={k} m={m} p={p}: kernels {r['device_ms']:.3f} ms; ess {r['ess'].tolist()} "
                    f"rhat {r['rhat'].tolist()}\n")
