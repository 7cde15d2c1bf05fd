"""The user-facing mirror of the reference's R API (bootstrap_filter, auxiliary_filter, resample_move_filter,
pmmh, resample_*_cpp) on the GPU: argument contract and return objects (R/particle_filter_core.R:248-266,
R/pmmh.R:599-608), plus the reference's own end-to-end statistical assertions."""
import numpy as np
import pytest

import bayesssm_b200 as b

pytestmark = pytest.mark.gpu


def _nonlinear_data(T, rng, phi=0.7, sx=1.0, sy=0.5):
    x = np.zeros(T + 1)
    x[0] = rng.standard_normal()
    for t in range(1, T + 1):
        x[t] = phi * x[t - 1] + np.sin(x[t - 1]) + sx * rng.standard_normal()
    return x, x[1:] + sy * rng.standard_normal(T)


def test_bootstrap_filter_return_object_and_rmse():
    # tests/testthat/test-bootstrap_filter.R:115-207
    rng = np.random.default_rng(1405)
    x, y = _nonlinear_data(50, rng)
    m = b.models.nonlinear_ar()
    r = b.bootstrap_filter(y, 100, m.init_fn, m.transition_fn, m.log_likelihood_fn, resample_algorithm="SISAR",
                           resample_fn="systematic", seed=1405, phi=0.7, sigma_x=1.0, sigma_y=0.5)
    assert {"state_est", "ess", "loglike", "loglike_history", "algorithm", "resample_algorithm",
            "particles_history", "weights_history"} <= set(r)
    assert r["algorithm"] == "BPF" and r["resample_algorithm"] == "SISAR"
    assert r["state_est"].shape == (51,) and r["ess"].shape == (51,) and r["loglike_history"].shape == (50,)
    assert r["particles_history"].shape == (51, 100) and r["weights_history"].shape == (51, 100)
    np.testing.assert_allclose(r["weights_history"].sum(axis=1), 1.0, rtol=1e-12)
    assert np.sqrt(np.mean((r["state_est"][1:] - x[1:]) ** 2)) < 0.6
    r2 = b.bootstrap_filter(y, 100, m.init_fn, m.transition_fn, m.log_likelihood_fn, return_particles=False,
                            seed=1405, phi=0.7, sigma_x=1.0, sigma_y=0.5)
    assert "particles_history" not in r2 and "weights_history" not in r2


def test_two_dimensional_particles():
    # tests/testthat/test-bootstrap_filter.R:211-230
    m = b.models.random_walk_2d()
    r = b.bootstrap_filter(np.zeros(10), 64, m.init_fn, m.transition_fn, m.log_likelihood_fn, seed=1, phi=0.1)
    assert r["state_est"].shape == (11, 2) and r["particles_history"].shape == (11, 128)


def test_apf_and_rmpf_beat_bpf():
    # tests/testthat/test-auxiliary_filter.R:1-54, test-resample_move_filter.R:1-62 (N = 20; replicated to de-noise)
    rng = np.random.default_rng(1405)
    m = b.models.random_walk_drift()
    wins_apf = wins_rm = 0
    for rep in range(40):
        x = np.cumsum(1.0 + rng.standard_normal(50))
        y = x + 0.1 * rng.standard_normal(50)
        kw = dict(seed=100 + rep, mu=1.0, sigma=0.1, return_particles=False)
        bp = b.bootstrap_filter(y, 20, m.init_fn, m.transition_fn, m.log_likelihood_fn, **kw)
        ap = b.auxiliary_filter(y, 20, m.init_fn, m.transition_fn, m.log_likelihood_fn, m.aux_log_likelihood_fn, **kw)
        rm = b.resample_move_filter(y, 20, m.init_fn, m.transition_fn, m.log_likelihood_fn, m.move_fn, **kw)
        assert ap["algorithm"] == "APF" and rm["algorithm"] == "RMPF" and rm["resample_algorithm"] == "SISR"
        mse = lambda r: np.mean((r["state_est"][1:] - x) ** 2)
        wins_apf += mse(ap) < mse(bp)
        wins_rm += mse(rm) < mse(bp)
    assert wins_apf >= 24 and wins_rm >= 20


def test_resample_cpp_wrappers():
    # tests/testthat/test-resampling.R:71-102,204-220
    w = np.array([0.1, 0.2, 0.3, 0.2, 0.2])
    for fn in (b.resample_multinomial_cpp, b.resample_stratified_cpp, b.resample_systematic_cpp):
        idx = fn(5, w, rng=1)
        assert idx.dtype == np.int32 and idx.shape == (5,) and idx.min() >= 1 and idx.max() <= 5
        with pytest.raises(ValueError, match="Weights must be non-negative"):
            fn(3, [-1, 1, 2])
        with pytest.raises(ValueError, match="Sum of weights must be greater than 0"):
            fn(3, [0, 0, 0])
    parts = np.arange(1, 7).reshape(2, 3).T
    for fn in (b.resample_multinomial, b.resample_stratified, b.resample_systematic):
        out = fn(parts, np.ones(3) / 3, rng=2)
        assert out.shape == parts.shape and np.isin(out, parts).all()
        with pytest.raises(ValueError, match="must match"):
            fn(np.arange(4), np.ones(3) / 3)


def test_pmmh_output_object_and_dead_arguments():
    # README.md:153-195 configuration (config C1), shortened; tests/testthat/test-pmmh.R:404-466: the resample_*
    # arguments of pmmh() are dead, so passing them must not change the chains
    rng = np.random.default_rng(1405)
    _, y = _nonlinear_data(20, rng, phi=0.8)
    m = b.models.nonlinear_ar()
    pri = {"phi": b.priors.uniform(0, 1), "sigma_x": b.priors.exponential(1), "sigma_y": b.priors.exponential(1)}
    init = [{"phi": 0.8, "sigma_x": 1.0, "sigma_y": 0.5}, {"phi": 0.5, "sigma_x": 0.5, "sigma_y": 1.0}]
    kw = dict(y=y, m=300, init_fn=m.init_fn, transition_fn=m.transition_fn, log_likelihood_fn=m.log_likelihood_fn,
              log_priors=pri, pilot_init_params=init, burn_in=50, num_chains=2,
              param_transform={"phi": "logit", "sigma_x": "log", "sigma_y": "log"},
              tune_control=b.default_tune_control(pilot_m=200, pilot_n=100, pilot_reps=10), seed=1405, print_result=False)
    with pytest.warns(UserWarning):
        out = b.pmmh(b.bootstrap_filter, **kw)
    tc = out["theta_chain"]
    assert list(tc.columns) == ["chain", "phi", "sigma_x", "sigma_y"]
    assert all(isinstance(v, str) for v in tc["chain"])      # character ids, as bind_rows(.id = "chain")
    assert len(tc) == 2 * (300 - 50) and set(tc["chain"]) == {"1", "2"}
    assert set(out["diagnostics"]) == {"ess", "rhat"} and set(out["diagnostics"]["ess"]) == {"phi", "sigma_x", "sigma_y"}
    assert ((tc["phi"] > 0) & (tc["phi"] < 1)).all() and (tc["sigma_x"] > 0).all()
    assert (50 <= out["target_n"]).all() and (out["target_n"] <= 1000).all()
    with pytest.warns(UserWarning):
        out2 = b.pmmh(b.bootstrap_filter, resample_algorithm="SISR", resample_fn="systematic", **kw)
    assert out2["theta_chain"].equals(tc)
    with pytest.raises(ValueError, match="Initial parameter values are invalid"):
        b.pmmh(b.bootstrap_filter, **{**kw, "pilot_init_params": [{"phi": 1.5, "sigma_x": 1.0, "sigma_y": 0.5}] * 2})


def test_pmmh_posterior_recovers_parameters():
    # tests/testthat/test-pmmh_tuning.R:505-575 in spirit: posterior mean within rel-tol 0.5 of the truth
    rng = np.random.default_rng(7)
    m = b.models.linear_gaussian()
    x, ys = rng.standard_normal(), []
    for _ in range(100):
        x = 0.8 * x + rng.standard_normal()
        ys.append(x + rng.standard_normal())
    pri = {"phi": b.priors.uniform(0, 1), "sigma_x": b.priors.exponential(1), "sigma_y": b.priors.exponential(1)}
    init = [{"phi": 0.5, "sigma_x": 1.0, "sigma_y": 1.0}] * 4
    with pytest.warns(UserWarning):
        out = b.pmmh(b.bootstrap_filter, np.array(ys), 1500, m.init_fn, m.transition_fn, m.log_likelihood_fn, pri, init,
                     burn_in=300, num_chains=4, param_transform={"phi": "logit", "sigma_x": "log", "sigma_y": "log"},
                     tune_control=b.default_tune_control(pilot_m=400, pilot_reps=20), seed=3, print_result=False,
                     precision="f32")
    tc = out["theta_chain"]
    # sigma_x and sigma_y are only jointly identified at T = 100 (their squares trade off): check phi, sigma_x and
    # the total innovation scale
    tot = np.sqrt(tc["sigma_x"] ** 2 + tc["sigma_y"] ** 2).mean()
    assert abs(tc["phi"].mean() - 0.8) < 0.4 and abs(tc["sigma_x"].mean() - 1.0) < 0.5 and abs(tot - np.sqrt(2.0)) < 0.5


def test_pmmh_return_latent_state_est(orc, engine):
    """R/pmmh.R:420,494-499,604-606: the state estimate of the filter run behind every draw travels with the chain.
    Checked against oracle filter runs at the accepted draws (same Philox ids: run_id = main phase << 28 | iteration,
    stream = global chain id), and carried over unchanged on rejections."""
    import bayesssm_b200 as b
    from bayesssm_b200.pmmh import default_tune_control, run_chains
    from bayesssm_b200 import _native as nat
    rng = np.random.default_rng(1405)
    x, ys = rng.standard_normal(), []
    for _ in range(10):
        x = 0.8 * x + np.sin(x) + rng.standard_normal()
        ys.append(x + 0.5 * rng.standard_normal())
    y = np.array(ys)
    mdl = b.models.nonlinear_ar()
    pri = [b.priors.uniform(0, 1), b.priors.exponential(1), b.priors.exponential(1)]
    init = np.array([[0.8, 1.0, 0.5], [0.6, 0.8, 0.7]])
    chol = np.tile(np.diag([0.3, 0.2, 0.2]), (2, 1, 1))
    m, N, seed = 25, 300, 11
    out = run_chains(engine, mdl, nat.BPF, y, init, pri, [nat.TR_LOGIT, nat.TR_LOG, nat.TR_LOG], default_tune_control(), m,
                     seed, chain_id_base=7, fixed_num_particles=N, precision=nat.F64, skip_pilot=True, proposal_chol=chol,
                     return_latent_state_est=True)
    lat, th = out["latent_state_chain"], out["theta_chain"]
    assert lat.shape == (2, m, len(y) + 1, 1)
    moved = 0
    for c in range(2):
        for i in range(m):
            if i > 0 and np.array_equal(th[c, i], th[c, i - 1]):
                np.testing.assert_array_equal(lat[c, i], lat[c, i - 1])       # rejected: carried over
                continue
            ref = orc.particle_filter(0, 0, 2, 0, N, y, th[c, i], seed=seed, run_id=(3 << 28) | i, stream=7 + c)
            np.testing.assert_allclose(lat[c, i, :, 0], ref["state_est"][:, 0], rtol=1e-9, atol=1e-12)
            moved += 1
    assert moved > 4
    # user-facing form: a list per chain of post-burn-in vectors
    res = b.pmmh(b.bootstrap_filter, y, m=12, init_fn=mdl.init_fn, transition_fn=mdl.transition_fn,
                 log_likelihood_fn=mdl.log_likelihood_fn,
                 log_priors={"phi": pri[0], "sigma_x": pri[1], "sigma_y": pri[2]},
                 pilot_init_params=[{"phi": .8, "sigma_x": 1., "sigma_y": .5}] * 2, burn_in=2, num_chains=2,
                 tune_control=dict(default_tune_control(), pilot_m=20, pilot_n=50, pilot_reps=4),
                 return_latent_state_est=True, seed=5, ctx=engine, print_result=False)
    assert len(res["latent_state_chain"]) == 2 and len(res["latent_state_chain"][0]) == 10
    assert res["latent_state_chain"][0][0].shape == (len(y) + 1,)


def test_replicate_filters_sharded_is_placement_independent(engine):
    """Config C3 across GPUs: filters are split by global filter id (= Philox stream); the shard [base, base+count)
    run alone reproduces the same rows of the full batch (world = 1 here; the gather is exercised on CPU with gloo)."""
    from bayesssm_b200 import distributed as D
    mdl = b.models.linear_gaussian()
    rng = np.random.default_rng(2)
    y = rng.standard_normal(30)
    kw = dict(resample_algorithm="SISR", precision="f64", seed=9, phi=0.8, sigma_x=1.0, sigma_y=1.0)
    full = D.replicate_filters_sharded(y, 2000, mdl, 12, 0, 1, ctx=engine, **kw)
    assert full["loglike"].shape == (12,) and len(np.unique(full["loglike"])) == 12
    from bayesssm_b200.filters import _particle_filter_core
    base, count = D.shard_chains(12, 1, 3)
    part = _particle_filter_core(y, 2000, mdl, "BPF", "SISR", "stratified", None, False, None,
                                 dict(phi=0.8, sigma_x=1.0, sigma_y=1.0), "f64", 9, engine, num_filters=count, stream_base=base)
    np.testing.assert_array_equal(part["loglike"], full["loglike"][base:base + count])


def test_exact_gillespie_sir_through_the_public_api():
    # vignettes/articles/stochastic-sir-model.Rmd:128-176, 285-309: data from the exact jump process (numpy restatement of
    # simulate_epidemic), filtered with the same model on the device; pmmh() runs on it as on any built-in model
    rng = np.random.default_rng(2025)
    s, i, latent = 430, 70, []
    for _ in range(10):
        t = 0.0
        while i > 0:
            ri, rr = 0.5 / 500 * s * i, 0.2 * i
            dt = rng.exponential(1.0 / (ri + rr))
            if t + dt > 1.0:
                break
            t += dt
            if rng.random() < ri / (ri + rr):
                s, i = s - 1, i + 1
            else:
                i -= 1
        latent.append(i)
    latent = np.array(latent, dtype=float)
    y = rng.poisson(latent).astype(float)
    m = b.models.sir_gillespie()
    kw = dict(seed=7, pop=500.0, I0=70.0, return_particles=False)
    r = b.bootstrap_filter(y, 2000, m.init_fn, m.transition_fn, m.log_likelihood_fn, **{"lambda": 0.5, "gamma": 0.2}, **kw)
    assert r["state_est"].shape == (11, 2)
    assert np.sqrt(np.mean((r["state_est"][1:, 1] - latent) ** 2)) < 12.0          # Poisson noise of ~ sqrt(150) per observation
    assert np.all(r["state_est"][:, 0] + r["state_est"][:, 1] <= 500.0 + 1e-9)
    # a far-off infection rate explains the data much worse
    r_bad = b.bootstrap_filter(y, 2000, m.init_fn, m.transition_fn, m.log_likelihood_fn, **{"lambda": 1.5, "gamma": 0.2}, **kw)
    assert r["loglike"] > r_bad["loglike"] + 10
    out = b.pmmh(b.bootstrap_filter, y, m=150, init_fn=m.init_fn, transition_fn=m.transition_fn, log_likelihood_fn=m.log_likelihood_fn,
                 log_priors={"lambda": b.priors.half_normal(1.0), "gamma": b.priors.half_normal(2.0)},
                 pilot_init_params=[{"lambda": 0.6, "gamma": 0.3}, {"lambda": 0.4, "gamma": 0.15}], burn_in=50, num_chains=2,
                 param_transform={"lambda": "log", "gamma": "log"}, seed=1405, pop=500.0, I0=70.0,
                 tune_control=b.default_tune_control(pilot_m=50, pilot_n=100, pilot_reps=4), verbose=False)
    tc = out["theta_chain"]
    assert 0.2 < tc["lambda"].mean() < 1.0 and 0.05 < tc["gamma"].mean() < 0.6
