"""GPU parity of the particle filters (general engine, through the C ABI) against the oracle restating
R/particle_filter_core.R, with identical injected particles and noise: log-likelihood within 1e-6 relative
(north star; measured ~1e-13), identical ancestors, state estimates / ESS to 1e-9."""
import numpy as np
import pytest

import engine_helpers as eh
from bayesssm_b200 import _native as nat

pytestmark = pytest.mark.gpu

AR, LG, RWD, SIR, ARCOS, RW2D = range(6)
THETA = {AR: [0.8, 1.0, 0.5], LG: [0.8, 1.0, 1.0], RWD: [1.0, 0.3], SIR: [0.5, 0.2, 500.0, 70.0], ARCOS: [0.8, 1.0, 0.5], RW2D: [0.1]}


def sim_y(model, T, rng):
    if model == SIR:
        return rng.poisson(60, T).astype(float)
    if model == RWD:
        return np.cumsum(1.0 + rng.standard_normal(T)) + 0.3 * rng.standard_normal(T)
    x, ys = rng.standard_normal(), []
    for _ in range(T):
        x = 0.8 * x + (np.sin(x) if model in (AR, ARCOS) else 0.0) + rng.standard_normal()
        ys.append((np.cos(x) if model == ARCOS else x) + 0.5 * rng.standard_normal())
    return np.array(ys)


def compare(orc, engine, model, algorithm, ralg, rfn, N, T, obs_times=None, threshold=-1.0, seed=0, hist=True):
    rng = np.random.default_rng(1000 + 17 * model + 5 * algorithm + ralg + 3 * rfn + N)
    y = sim_y(model, T, rng)
    n_time = int(obs_times[-1]) if obs_times is not None else T
    noise = orc.make_noise(model, N, T, n_time, rng)
    th = THETA[model]
    ref = orc.particle_filter(model, algorithm, ralg, rfn, N, y, th, threshold=threshold, obs_times=obs_times,
                              noise=noise, return_particles=hist, want_ancestors=True)
    got = eh.filter_run(engine, model, algorithm, ralg, rfn, N, y, th, threshold=threshold, obs_times=obs_times,
                        noise=noise, precision=nat.F64, return_particles=hist, want_ancestors=True)
    assert ref["status"] == 0 and got["status"][0] == 0
    assert got["early_exit"][0] == ref["early_exit"]
    assert got["n_resampled"][0] == ref["n_resampled"]
    # north-star tolerance: 1e-6 relative on the log-likelihood; we hold a far tighter one
    assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-10 * max(1.0, abs(ref["loglike"]))
    np.testing.assert_allclose(got["loglike_history"][0], ref["loglike_history"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-9)
    np.testing.assert_allclose(got["state_est"][0], ref["state_est"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(got["ancestors_history"][0], ref["ancestors_history"])
    if algorithm == 1:
        assert np.array_equal(got["ancestors_aux_history"][0], ref["ancestors_aux_history"])
    if hist:
        np.testing.assert_allclose(got["particles_history"][0], ref["particles_history"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(got["weights_history"][0], ref["weights_history"], rtol=1e-9, atol=1e-300)
    return ref, got


@pytest.mark.parametrize("rfn", [0, 1, 2])
@pytest.mark.parametrize("ralg", [0, 1, 2])
def test_bpf_nonlinear_ar(orc, engine, ralg, rfn):
    compare(orc, engine, AR, 0, ralg, rfn, N=1000, T=20)      # config C1 shape: README model, T=20, N=1000


@pytest.mark.parametrize("model", [LG, RWD, SIR, ARCOS, RW2D])
def test_bpf_other_models(orc, engine, model):
    compare(orc, engine, model, 0, 2, 0, N=777, T=15)


@pytest.mark.parametrize("model", [AR, RWD, SIR, LG])
@pytest.mark.parametrize("rfn", [0, 1])
def test_apf(orc, engine, model, rfn):
    compare(orc, engine, model, 1, 2, rfn, N=500, T=12)


@pytest.mark.parametrize("model", [RWD, SIR, AR])
def test_rmpf(orc, engine, model):
    compare(orc, engine, model, 2, 2, 0, N=500, T=12)


def test_obs_times_gaps(orc, engine):
    compare(orc, engine, AR, 0, 2, 0, N=300, T=6, obs_times=[1, 2, 4, 7, 8, 12])


def test_explicit_threshold(orc, engine):
    compare(orc, engine, AR, 0, 2, 1, N=4096, T=10, threshold=0.9 * 4096)


def test_large_n_single_step_sizes(orc, engine):
    compare(orc, engine, AR, 0, 1, 0, N=70001, T=3, hist=False)


def test_early_exit(orc, engine):
    # all log-weights < -1e8 => loglike = -Inf and a truncated, zero-filled result (R/particle_filter_core.R:189-202)
    y = np.array([0.1, 1e6, 0.2])
    rng = np.random.default_rng(1)
    noise = orc.make_noise(AR, 64, 3, 3, rng)
    th = [0.8, 1.0, 1e-3]
    ref = orc.particle_filter(AR, 0, 2, 0, 64, y, th, noise=noise)
    got = eh.filter_run(engine, AR, 0, 2, 0, 64, y, th, noise=noise)
    assert ref["early_exit"] == 1 and got["early_exit"][0] == 1
    assert got["loglike"][0] == -np.inf and ref["loglike"] == -np.inf
    np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-9)
    np.testing.assert_allclose(got["state_est"][0], ref["state_est"], rtol=1e-9, atol=1e-12)


def test_philox_mode_matches_oracle_philox(orc, engine):
    # same counter-based generator on both sides: no injected buffers at all
    rng = np.random.default_rng(2)
    y = sim_y(AR, 25, rng)
    for stream in (0, 5):
        ref = orc.particle_filter(AR, 0, 2, 0, 2000, y, THETA[AR], seed=1405, run_id=3, stream=stream)
        got = eh.filter_run(engine, AR, 0, 2, 0, 2000, y, THETA[AR], seed=1405, run_id=3, stream_base=stream)
        assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
        np.testing.assert_allclose(got["state_est"][0], ref["state_est"], rtol=1e-6, atol=1e-6)


def test_batched_filters_are_independent_streams(orc, engine):
    rng = np.random.default_rng(3)
    y = sim_y(LG, 30, rng)
    got = eh.filter_run(engine, LG, 0, 1, 0, 512, y, THETA[LG], seed=7, num_filters=6)
    for c in (0, 3, 5):
        ref = orc.particle_filter(LG, 0, 1, 0, 512, y, THETA[LG], seed=7, stream=c)
        assert abs(got["loglike"][c] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])


def test_f32_precision_statistically_consistent(orc, engine):
    # throughput precision (fp32 state / weights, fp64 cdf and accumulators) against the exact Kalman value
    rng = np.random.default_rng(4)
    y = sim_y(LG, 100, rng)
    exact = orc.kalman_loglik(y, 0.8, 1.0, 1.0)
    got = eh.filter_run(engine, LG, 0, 1, 0, 4096, y, THETA[LG], seed=11, num_filters=32, precision=nat.F32)
    lls = got["loglike"]
    est = np.log(np.mean(np.exp(lls - lls.max()))) + lls.max()
    se = lls.std(ddof=1) / np.sqrt(len(lls))
    assert abs(est - exact) < 3 * se + 0.02, (est, exact, se)


def test_more_than_65535_filters_in_one_batch(engine):
    # .pilot_run of 1024 chains x 100 replicates is one batch of 102 400 filters: the filter index lives in grid.x
    rng = np.random.default_rng(13)
    y = sim_y(LG, 4, rng)
    got = eh.filter_run(engine, LG, 0, 2, 0, 64, y, THETA[LG], seed=3, num_filters=70000, precision=nat.F64,
                        engine=nat.ENGINE_GENERAL)
    assert (got["status"] == 0).all() and np.isfinite(got["loglike"]).all()
    assert len(np.unique(got["loglike"])) > 69000


def test_histories_stream_through_a_two_row_ring(orc, engine):
    # return_particles: the device holds two history rows at a time; rows travel to the caller's buffers on a copy stream while
    # the filter runs (R/particle_filter_core.R:100-116,242-264).  A batch of three filters over 40 observations -- the ring
    # wraps 20 times -- one of which stops early (its remaining rows stay zero, :189-202)
    rng = np.random.default_rng(12)
    T, N = 40, 3000
    y = sim_y(AR, T, rng)
    thetas = np.array([THETA[AR], [0.7, 1.1, 0.6], THETA[AR]])
    got = eh.filter_run(engine, AR, 0, 2, 0, N, y, thetas, seed=8, precision=nat.F64, return_particles=True)
    for c in range(3):
        ref = orc.particle_filter(AR, 0, 2, 0, N, y, thetas[c], seed=8, stream=c, return_particles=True)
        np.testing.assert_allclose(got["particles_history"][c], ref["particles_history"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got["weights_history"][c], ref["weights_history"], rtol=1e-9, atol=1e-15)
    y2 = y.copy()
    y2[25] = 1e6                                  # every log-weight below -1e8 at observation 25 with sigma_y = 1e-3
    th = [0.8, 1.0, 1e-3]
    got = eh.filter_run(engine, AR, 0, 1, 0, 512, y2, th, seed=8, precision=nat.F64, return_particles=True)
    assert got["early_exit"][0] == 1
    ph, wh = got["particles_history"][0], got["weights_history"][0]
    assert np.abs(ph[:26]).max() > 0 and not ph[26:].any() and not wh[26:].any()


# ---- stochastic SIR with the exact (Gillespie) daily step: uniforms drawn on demand from the particle's Philox stream ----
GIL = 6
THETA_GIL = [0.5, 0.2, 500.0, 70.0]


@pytest.mark.parametrize("algorithm,ralg", [(0, 2), (0, 1), (1, 2), (2, 2)])
def test_gillespie_sir_matches_oracle_philox(orc, engine, algorithm, ralg):
    # vignettes/articles/stochastic-sir-model.Rmd:152-176 (epidemic_step) on both sides, the same Philox words: the event
    # sequences are identical, so the parity is that of the other Philox-mode tests
    y = np.array([82, 95, 118, 130, 151, 160, 158, 149], dtype=float)
    for stream in (0, 3):
        ref = orc.particle_filter(GIL, algorithm, ralg, 0, 1500, y, THETA_GIL, seed=11, run_id=1, stream=stream, return_particles=True)
        got = eh.filter_run(engine, GIL, algorithm, ralg, 0, 1500, y, THETA_GIL, seed=11, run_id=1, stream_base=stream,
                            precision=nat.F64, return_particles=True)
        assert ref["status"] == 0 and got["status"][0] == 0
        assert got["n_resampled"][0] == ref["n_resampled"]
        assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-9 * abs(ref["loglike"])
        np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-9)
        np.testing.assert_allclose(got["state_est"][0], ref["state_est"], rtol=1e-9, atol=1e-9)
        np.testing.assert_array_equal(got["particles_history"][0], ref["particles_history"])    # integer states: exactly
    # throughput precision: the same integer trajectories (the step itself is always fp64), weights in fp32
    got32 = eh.filter_run(engine, GIL, 0, 2, 0, 1500, y, THETA_GIL, seed=11, run_id=1, stream_base=3, precision=nat.F32)
    ref = orc.particle_filter(GIL, 0, 2, 0, 1500, y, THETA_GIL, seed=11, run_id=1, stream=3)
    assert abs(got32["loglike"][0] - ref["loglike"]) < 0.05


def test_gillespie_sir_rejects_injected_noise(orc, engine):
    y = np.array([80.0, 90.0])
    noise = orc.make_noise(SIR, 64, 2, 2, np.random.default_rng(0))
    with pytest.raises(Exception):
        eh.filter_run(engine, GIL, 0, 2, 0, 64, y, THETA_GIL, noise=noise)


# ---- carried-weights mode (bssm_filter_config::carry_weights): a stated deviation from the reference's weight rule ----
@pytest.mark.parametrize("model,algorithm,ralg", [(LG, 0, 0), (AR, 0, 2), (SIR, 0, 2), (RWD, 2, 1)])
def test_carried_weights_match_the_oracle(orc, engine, model, algorithm, ralg):
    rng = np.random.default_rng(50 + model)
    T, N = 9, 1200
    y = sim_y(model, T, rng)
    noise = orc.make_noise(model, N, T, T, rng)
    ref = orc.particle_filter(model, algorithm, ralg, 0, N, y, THETA[model], noise=noise, return_particles=True, carry_weights=True)
    got = eh.filter_run(engine, model, algorithm, ralg, 0, N, y, THETA[model], noise=noise, precision=nat.F64, return_particles=True,
                        carry_weights=True)
    assert ref["status"] == 0 and got["status"][0] == 0 and got["n_resampled"][0] == ref["n_resampled"]
    assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-9 * max(1.0, abs(ref["loglike"]))
    np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-8)
    np.testing.assert_allclose(got["state_est"][0], ref["state_est"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(got["weights_history"][0], ref["weights_history"], rtol=1e-8, atol=1e-300)
    if ralg != 1 and algorithm == 0:
        plain = eh.filter_run(engine, model, algorithm, ralg, 0, N, y, THETA[model], noise=noise, precision=nat.F64)
        assert plain["loglike"][0] != got["loglike"][0]


def test_carried_weights_are_refused_where_they_are_not_served(engine):
    y = sim_y(AR, 5, np.random.default_rng(1))
    with pytest.raises(Exception, match="carry_weights"):
        eh.filter_run(engine, AR, 1, 2, 0, 500, y, THETA[AR], carry_weights=True)                      # auxiliary filter
    for eng in (nat.ENGINE_PERSISTENT, nat.ENGINE_STREAM):
        with pytest.raises(Exception):
            eh.filter_run(engine, AR, 0, 2, 0, 5000, y, THETA[AR], precision=nat.F32, engine=eng, carry_weights=True)
    # AUTO in the throughput precision: served by the general kernels
    a = eh.filter_run(engine, AR, 0, 2, 0, 50000, y, THETA[AR], seed=4, precision=nat.F32, carry_weights=True)
    b = eh.filter_run(engine, AR, 0, 2, 0, 50000, y, THETA[AR], seed=4, precision=nat.F32, carry_weights=True, engine=nat.ENGINE_GENERAL)
    assert a["status"][0] == 0 and a["loglike"][0] == b["loglike"][0]
