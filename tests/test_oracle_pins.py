"""Pins the CPU oracle against every known-answer property the reference's own tests hold for the
hot path.  No GPU."""
import numpy as np
import pytest


def test_validation_errors(orc):
    # tests/testthat/test-resampling.R:2-28
    for kind in ("multinomial", "multinomial_rcpp", "stratified", "systematic"):
        with pytest.raises(orc.OracleError, match="Weights must be non-negative"):
            orc.resample(kind, [-1, 1, 2], [0.5, 0.5, 0.5])
        with pytest.raises(orc.OracleError, match="Sum of weights must be greater than 0"):
            orc.resample(kind, [0, 0, 0], [0.5, 0.5, 0.5])


def test_proportions(orc):
    # tests/testthat/test-resampling.R:29-47 (10 000 replicates, tolerance 0.05)
    rng = np.random.default_rng(1405)
    w = np.array([0.1, 0.2, 0.3, 0.2, 0.2])
    for kind in ("multinomial", "multinomial_rcpp", "stratified", "systematic"):
        counts = np.zeros(5)
        for _ in range(10000):
            idx = orc.resample(kind, w, rng.random(5))
            counts += np.bincount(idx - 1, minlength=5)
        np.testing.assert_allclose(counts / 50000, w, atol=0.05 * w.max())


def test_structural_known_answers(orc):
    # tests/testthat/test-resampling.R:48-68
    rng = np.random.default_rng(7)
    w = np.array([0.1, 0.5, 0.1, 0.15, 0.15])
    for _ in range(2000):
        s = orc.resample("stratified", w, rng.random(5))
        assert s[1] == 2 and s[2] == 2
        y = orc.resample("systematic", w, rng.random(1))
        assert y[1] == 2 and y[2] == 2
        if y[0] == 1:
            assert y[3] == 3
        if y[0] == 2:
            assert y[3] == 4


def test_degenerate_weights(orc):
    # tests/testthat/test-resampling.R:190-202
    rng = np.random.default_rng(123)
    w = [0, 0, 1, 0, 0]
    for kind in ("multinomial", "multinomial_rcpp", "stratified", "systematic"):
        assert (orc.resample(kind, w, rng.random(5)) == 3).all()


def test_tie_rule(orc):
    # pos == c[j] selects j (uniform weights n=4, U=0 => 1,1,2,3)
    assert orc.resample("systematic", [0.25] * 4, [0.0]).tolist() == [1, 1, 2, 3]
    assert orc.resample("stratified", [0.25] * 4, [0.0] * 4).tolist() == [1, 1, 2, 3]


def test_rcpp_walker_branch_distribution(orc):
    # Rcpp::sample switches to Walker alias tables above 200 "large" probabilities
    rng = np.random.default_rng(3)
    n = 400
    w = rng.random(n) + 0.5
    counts = np.zeros(n)
    for _ in range(300):
        counts += np.bincount(orc.resample("multinomial_rcpp", w, rng.random(n)) - 1, minlength=n)
    p = w / w.sum()
    z = (counts - 300 * n * p) / np.sqrt(300 * n * p * (1 - p))
    assert np.abs(z).max() < 5.0


def test_transforms_and_jacobian(orc):
    # tests/testthat/test-utils.R:26-59 (TR_LOGIT = 2, TR_LOG = 1)
    th = 0.5
    assert orc.transform(th, 2) == 0.0
    assert orc.transform(th, 2) == np.log(th / (1 - th))
    assert orc.back_transform(np.log(th / (1 - th)), 2) == th
    assert orc.log_jacobian([th], [2]) == pytest.approx(-np.log(th * (1 - th)), abs=1e-15)
    assert orc.log_jacobian([2.0], [1]) == pytest.approx(np.log(2.0))
    assert orc.log_jacobian([2.0], [0]) == 0.0


def test_priors_match_r_densities(orc):
    from scipy import stats
    assert orc.log_prior(1, 0.0, 1.0, 0.3) == pytest.approx(stats.norm.logpdf(0.3), abs=1e-14)
    assert orc.log_prior(2, 1.0, 0.0, 0.7) == pytest.approx(stats.expon.logpdf(0.7), abs=1e-14)
    assert orc.log_prior(2, 1.0, 0.0, -0.1) == -np.inf
    assert orc.log_prior(3, 0.0, 1.0, 0.5) == 0.0
    assert orc.log_prior(3, 0.0, 1.0, 1.5) == -np.inf
    assert orc.log_prior(4, 2.0, 0.0, 0.4) == pytest.approx(stats.halfnorm.logpdf(0.4, scale=2.0), abs=1e-14)


def test_bootstrap_filter_rmse(orc):
    # tests/testthat/test-bootstrap_filter.R:149-207: nonlinear AR, T=50, N=100, SISAR + systematic => RMSE < 0.5
    rng = np.random.default_rng(1405)
    T, phi, sx, sy = 50, 0.7, 1.0, 0.5   # sigma values as in the reference test's data simulation
    x = np.zeros(T + 1)
    x[0] = rng.standard_normal()
    for t in range(1, T + 1):
        x[t] = phi * x[t - 1] + np.sin(x[t - 1]) + sx * rng.standard_normal()
    y = x[1:] + sy * rng.standard_normal(T)
    r = orc.particle_filter(0, 0, 2, 1, 100, y, [phi, sx, sy], seed=1405)
    assert r["status"] == 0 and len(r["state_est"]) == T + 1
    rmse = np.sqrt(np.mean((r["state_est"][1:, 0] - x[1:]) ** 2))
    assert rmse < 0.6


def test_apf_and_rmpf_beat_bpf_on_drift_model(orc):
    # tests/testthat/test-auxiliary_filter.R:1-54 and test-resample_move_filter.R:1-62 (RW + drift, N=20)
    rng = np.random.default_rng(1405)
    T, mu = 50, 1.0
    wins_apf = wins_rmpf = 0
    for rep in range(40):
        x = np.cumsum(mu + rng.standard_normal(T))
        sig = 0.1
        y = x + sig * rng.standard_normal(T)
        kw = dict(N=20, y=y, theta=[mu, sig], seed=100 + rep)
        bpf = orc.particle_filter(2, 0, 2, 0, **kw)
        apf = orc.particle_filter(2, 1, 2, 0, **kw)
        rmpf = orc.particle_filter(2, 2, 1, 0, **kw)
        mse = lambda r: np.mean((r["state_est"][1:, 0] - x) ** 2)
        wins_apf += mse(apf) < mse(bpf)
        wins_rmpf += mse(rmpf) < mse(bpf)
    assert wins_apf >= 24 and wins_rmpf >= 20


def test_kalman_agreement_lg(orc):
    # linear-Gaussian model: PF log-likelihood vs exact Kalman within 3 MC standard errors (north star)
    rng = np.random.default_rng(5)
    T, phi, sx, sy = 100, 0.8, 1.0, 1.0
    x = 0.0 + rng.standard_normal()
    ys = []
    for _ in range(T):
        x = phi * x + sx * rng.standard_normal()
        ys.append(x + sy * rng.standard_normal())
    y = np.array(ys)
    exact = orc.kalman_loglik(y, phi, sx, sy)
    lls = np.array([orc.particle_filter(1, 0, 1, 0, 2000, y, [phi, sx, sy], seed=9, stream=s)["loglike"] for s in range(24)])
    est = np.log(np.mean(np.exp(lls - lls.max()))) + lls.max()
    se = lls.std(ddof=1) / np.sqrt(len(lls))
    assert abs(est - exact) < 3 * se + 0.02


def test_default_tune_control_clamp(orc):
    # R/pmmh_tuning.R:54-57: target_n clamped to [50, 1000]
    rng = np.random.default_rng(0)
    y = rng.standard_normal(10)
    r = orc.pmmh_chain(0, 0, y, [0.5, 1.0, 1.0], [3, 2, 2], [0, 1, 1], [1, 0, 0], [0, 1, 1], [0.1, 0.1, 0.1],
                       pilot_n=30, pilot_m=40, pilot_reps=5, m=20, chain_id=0, seed=1)
    assert r["status"] == 0 and 50 <= r["target_n"] <= 1000
    assert np.isfinite(r["theta_chain"]).all()


def test_pmmh_latent_state_chain_travels_with_the_draws(orc):
    # R/pmmh.R:400,420,494-499: current_state_est is replaced on acceptance only; draw i carries the state estimate of
    # the filter run (run id = main phase << 28 | iteration of the accepting proposal, stream = chain id) behind it
    rng = np.random.default_rng(1405)
    x, ys = rng.standard_normal(), []
    for _ in range(10):
        x = 0.8 * x + np.sin(x) + rng.standard_normal()
        ys.append(x + 0.5 * rng.standard_normal())
    y = np.array(ys)
    kw = dict(prior_kind=[3, 2, 2], prior_a=[0., 1., 1.], prior_b=[1., 0., 0.], transform=[2, 1, 1],
              pilot_proposal_sd=[0.1, 0.15, 0.2], pilot_n=64, pilot_m=20, pilot_reps=5, m=30, seed=7)
    r = orc.pmmh_chain(0, 0, y, [0.8, 1.0, 0.5], chain_id=3, return_latent_state_est=True, **kw)
    plain = orc.pmmh_chain(0, 0, y, [0.8, 1.0, 0.5], chain_id=3, **kw)
    np.testing.assert_array_equal(r["theta_chain"], plain["theta_chain"])     # asking for it changes nothing else
    np.testing.assert_array_equal(r["loglike_chain"], plain["loglike_chain"])
    lat, th = r["latent_state_chain"], r["theta_chain"]
    assert lat.shape == (30, len(y) + 1, 1)
    moved = stayed = 0
    for i in range(30):
        if i > 0 and np.array_equal(th[i], th[i - 1]):
            np.testing.assert_array_equal(lat[i], lat[i - 1])
            stayed += 1
            continue
        ref = orc.particle_filter(0, 0, 2, 0, r["target_n"], y, th[i], seed=7, run_id=(3 << 28) | i, stream=3)
        np.testing.assert_array_equal(lat[i, :, 0], ref["state_est"][:, 0])
        assert ref["loglike"] == r["loglike_chain"][i]
        moved += 1
    assert moved == r["n_accept"] + 1 and stayed > 0


def test_multinomial_by_sorted_uniforms_has_the_multinomial_law(orc):
    # orc_resample_multinomial_sorted (what the streaming engine's multinomial path is checked against): same offspring
    # proportions as the reference's test asks of resample_multinomial_cpp (tests/testthat/test-resampling.R:29-47), sorted
    # ancestors, degenerate weights (:190-202), and the reference's error strings through the shared cdf (:2-28)
    import ctypes as C
    L = orc.lib()
    L.orc_resample_multinomial_sorted.restype = C.c_int
    rng = np.random.default_rng(1405)
    w = np.array([0.1, 0.5, 0.1, 0.15, 0.15])
    counts = np.zeros(5)
    for _ in range(10000):
        u = np.ascontiguousarray(rng.random(6))
        idx = np.zeros(5, dtype=np.int32)
        assert L.orc_resample_multinomial_sorted(5, w.ctypes.data_as(C.POINTER(C.c_double)), u.ctypes.data_as(C.POINTER(C.c_double)),
                                                 idx.ctypes.data_as(C.POINTER(C.c_int32))) == 0
        assert (np.diff(idx) >= 0).all() and idx.min() >= 1 and idx.max() <= 5
        counts += np.bincount(idx - 1, minlength=5)
    assert np.abs(counts / counts.sum() - w).max() < 0.05 / 4
    # variance of one offspring count: n p (1 - p) (a stratified scheme would have far less)
    n, reps, p = 200, 4000, 0.3
    w2 = np.r_[p, np.full(n - 1, (1 - p) / (n - 1))]
    c0 = np.empty(reps)
    for r in range(reps):
        u = np.ascontiguousarray(rng.random(n + 1))
        idx = np.zeros(n, dtype=np.int32)
        L.orc_resample_multinomial_sorted(n, w2.ctypes.data_as(C.POINTER(C.c_double)), u.ctypes.data_as(C.POINTER(C.c_double)), idx.ctypes.data_as(C.POINTER(C.c_int32)))
        c0[r] = (idx == 1).sum()
    assert abs(c0.mean() - n * p) < 0.5 and abs(c0.var() / (n * p * (1 - p)) - 1) < 0.1
    w3 = np.array([0.0, 0.0, 1.0, 0.0, 0.0])
    idx = np.zeros(5, dtype=np.int32)
    u = np.ascontiguousarray(rng.random(6))
    L.orc_resample_multinomial_sorted(5, w3.ctypes.data_as(C.POINTER(C.c_double)), u.ctypes.data_as(C.POINTER(C.c_double)), idx.ctypes.data_as(C.POINTER(C.c_int32)))
    assert (idx == 3).all()


def test_gillespie_sir_step_has_the_law_of_the_reference_simulation(orc):
    """vignettes/articles/stochastic-sir-model.Rmd:152-176 (epidemic_step): one day of the exact SIR jump process from
    (s, i) = (430, 70), lambda = 0.5, gamma = 0.2, N = 500 -- the oracle's particles after one SIS step against an
    independent numpy simulation of the same process (means within 4 standard errors, spreads within 10 %)."""
    n = 20000
    out = orc.particle_filter(6, 0, 0, 0, n, np.array([75.0]), [0.5, 0.2, 500.0, 70.0], seed=5, return_particles=True)
    assert out["status"] == 0
    s1, i1 = out["particles_history"][1]
    assert np.all(s1 == np.round(s1)) and np.all(i1 == np.round(i1)) and np.all(s1 <= 430) and np.all(s1 + i1 <= 500) and np.all(i1 >= 0)
    rng = np.random.default_rng(1)
    sim = np.zeros((n, 2))
    for k in range(n):
        s, i, t = 430, 70, 0.0
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
        sim[k] = s, i
    for got, ref in ((s1, sim[:, 0]), (i1, sim[:, 1])):
        se = np.sqrt(got.var() / n + ref.var() / n)
        assert abs(got.mean() - ref.mean()) < 4 * se, (got.mean(), ref.mean(), se)
        assert abs(got.std() / ref.std() - 1) < 0.1
    # injected noise buffers cannot serve a data-dependent number of uniforms
    noise = orc.make_noise(3, 16, 1, 1, np.random.default_rng(0))
    assert orc.particle_filter(6, 0, 2, 0, 16, np.array([75.0]), [0.5, 0.2, 500.0, 70.0], noise=noise)["status"] != 0


def test_carried_weights_mode_is_standard_importance_sampling(orc):
    """The optional deviation that fixes SURVEY App. A1: with SIS (never resampling) and the weights carried from step to
    step, prod_t sum_i W_{t-1,i} g_t(x_i) is plain importance sampling of the likelihood -- it agrees with the exact Kalman
    value, while the reference's rule (each step's weights are its likelihoods only) gives the product of the marginal means."""
    rng = np.random.default_rng(3)
    x, ys = rng.standard_normal(), []
    for _ in range(6):
        x = 0.8 * x + rng.standard_normal()
        ys.append(x + rng.standard_normal())
    y, th = np.array(ys), [0.8, 1.0, 1.0]
    exact = orc.kalman_loglik(y, *th)
    carried = orc.particle_filter(1, 0, 0, 0, 400000, y, th, seed=1, carry_weights=True)
    plain = orc.particle_filter(1, 0, 0, 0, 400000, y, th, seed=1)
    assert abs(carried["loglike"] - exact) < 0.05 and abs(plain["loglike"] - exact) > 0.3
    assert carried["ess"][-1] < plain["ess"][-1]       # the carried weights degenerate, as they must
    # under SISR every step resamples: the two rules coincide
    a = orc.particle_filter(1, 0, 1, 0, 5000, y, th, seed=2, carry_weights=True)
    b = orc.particle_filter(1, 0, 1, 0, 5000, y, th, seed=2)
    assert a["loglike"] == b["loglike"]
    # not defined for the auxiliary filter's two-stage weights
    assert orc.particle_filter(1, 1, 2, 0, 100, y, th, seed=2, carry_weights=True)["status"] != 0
