"""BASELINE.json configurations at (or near) full size, checked through size-independent properties:
exact Kalman agreement (config 3), conservation / range invariants of the integer SIR state under APF and RMPF
(config 4), geometry independence and determinism of the throughput path (configs 2 and 5)."""
import numpy as np
import pytest

import engine_helpers as eh
from bayesssm_b200 import _native as nat
from test_filter_gpu import THETA, sim_y

pytestmark = pytest.mark.gpu
AR, LG, RWD, SIR = 0, 1, 2, 3


def test_config3_lg_256_filters_against_kalman(orc, engine):
    # linear-Gaussian SSM, T = 500, N = 2^16, 256 batched filters, SISR: log-mean-exp of the estimates within
    # 3 MC standard errors of the exact Kalman log-likelihood (north-star gate)
    rng = np.random.default_rng(3)
    y = sim_y(LG, 500, rng)
    exact = orc.kalman_loglik(y, 0.8, 1.0, 1.0)
    got = eh.filter_run(engine, LG, 0, 1, 0, 1 << 16, y, THETA[LG], seed=1405, num_filters=256, precision=nat.F32)
    assert (got["status"] == 0).all()
    lls = got["loglike"]
    est = np.log(np.mean(np.exp(lls - lls.max()))) + lls.max()
    se = lls.std(ddof=1) / np.sqrt(len(lls))
    assert abs(est - exact) < 3 * se + 5e-3, (est, exact, se)
    assert lls.std(ddof=1) < 0.2            # N = 65536: the estimator is tight
    np.testing.assert_allclose(got["ess"][:, 1:], 1 << 16)   # SISR: ESS reported as N after every resampling


@pytest.mark.parametrize("algorithm", [1, 2])
def test_config4_sir_apf_rmpf_n_2pow18(engine, algorithm):
    # stochastic SIR (chain-binomial), pop = 500, I0 = 70, lambda = 0.5, gamma = 0.2, Poisson observations,
    # T = 100, N = 2^18; invariants of every particle carry over to the weighted means: 0 <= S <= S0, 0 <= I, S + I <= pop
    rng = np.random.default_rng(4)
    S, I, ys = 430, 70, []
    for _ in range(100):
        ni = rng.binomial(S, 1 - np.exp(-0.5 * I / 500))
        nr = rng.binomial(I, 1 - np.exp(-0.2))
        S, I = S - ni, I + ni - nr
        ys.append(rng.poisson(max(I, 0)))
    y = np.array(ys, dtype=float)
    got = eh.filter_run(engine, SIR, algorithm, 2, 0, 1 << 18, y, [0.5, 0.2, 500.0, 70.0], seed=7, precision=nat.F64)
    assert got["status"][0] == 0 and np.isfinite(got["loglike"][0])
    se = got["state_est"][0]
    assert (se >= -1e-9).all() and (se.sum(axis=1) <= 500 + 1e-9).all()
    assert se[0, 0] == 430 and se[0, 1] == 70
    assert (se[:, 0] <= 430 + 1e-9).all()                        # particles never gain susceptibles
    assert np.abs(se[1:, 1] - y).mean() < 15.0                   # filtered infected count tracks the Poisson observations (sd ~ 12 at the peak)
    assert (got["ess"][0] <= (1 << 18) + 1e-6).all() and (got["ess"][0] > 1).all()


def test_config2_throughput_path_is_deterministic_and_geometry_free(engine, monkeypatch):
    # N = 2^20, one filter over the whole chip: repeated runs are bit-identical, and the f64 result does not
    # depend on how particles are split over CTAs (slice size) beyond summation order
    rng = np.random.default_rng(5)
    y = sim_y(AR, 12, rng)
    a = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=9, precision=nat.F32, engine=nat.ENGINE_PERSISTENT)
    b = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=9, precision=nat.F32, engine=nat.ENGINE_PERSISTENT)
    assert a["loglike"][0] == b["loglike"][0] and np.array_equal(a["state_est"], b["state_est"])
    ref = eh.filter_run(engine, AR, 0, 2, 0, 1 << 18, y, THETA[AR], seed=9, precision=nat.F64, engine=nat.ENGINE_PERSISTENT)
    monkeypatch.setenv("BSSM_FAST_NB", "1024")
    alt = eh.filter_run(engine, AR, 0, 2, 0, 1 << 18, y, THETA[AR], seed=9, precision=nat.F64, engine=nat.ENGINE_PERSISTENT)
    assert abs(ref["loglike"][0] - alt["loglike"][0]) <= 1e-9 * abs(ref["loglike"][0])
    np.testing.assert_allclose(ref["state_est"][0], alt["state_est"][0], rtol=1e-9, atol=1e-9)


def test_resampling_round_trip_properties_at_2pow20(engine):
    # every ancestor is a valid index; offspring counts follow the weights (|count_i - N w_i| < 1 for
    # stratified / systematic); ancestors are non-decreasing (the two-pointer sweep of src/resampling.cpp:32-37)
    rng = np.random.default_rng(6)
    n = 1 << 20
    w = np.exp(rng.standard_normal(n) * 2)
    for fn in ("stratified", "systematic"):
        idx = eh.resample(engine, fn, w, rng.random(1 if fn == "systematic" else n))
        assert idx.min() >= 1 and idx.max() <= n and (np.diff(idx) >= 0).all()
        counts = np.bincount(idx - 1, minlength=n)
        bound = 1.0 if fn == "systematic" else 2.0
        assert np.abs(counts - n * w / w.sum()).max() < bound + 1e-6
    idx = eh.resample(engine, "multinomial", w, rng.random(n))
    assert idx.min() >= 1 and idx.max() <= n
    counts = np.bincount(idx - 1, minlength=n)
    p = w / w.sum()
    top = np.argsort(p)[-50:]
    z = (counts[top] - n * p[top]) / np.sqrt(n * p[top])
    assert np.abs(z).max() < 6
