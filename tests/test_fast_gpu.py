"""The persistent bootstrap-filter kernel (bssm_fast.cuh) against the oracle and the general engine.
Same Philox streams on every side, so in f64 the persistent kernel reproduces the oracle's Philox-mode
filter up to summation order; in f32 (throughput precision) it is checked statistically (Kalman)."""
import numpy as np
import pytest

import engine_helpers as eh
from bayesssm_b200 import _native as nat
from test_filter_gpu import THETA, sim_y

pytestmark = pytest.mark.gpu
AR, LG, RWD = 0, 1, 2


@pytest.mark.parametrize("N,T", [(1000, 30), (4096, 40), (70001, 12), (1 << 17, 6)])
@pytest.mark.parametrize("rfn", [0, 1])
def test_f64_matches_oracle_philox(orc, engine, N, T, rfn):
    rng = np.random.default_rng(N + rfn)
    y = sim_y(AR, T, rng)
    ref = orc.particle_filter(AR, 0, 2, rfn, N, y, THETA[AR], seed=1405, run_id=2, stream=3)
    got = eh.filter_run(engine, AR, 0, 2, rfn, N, y, THETA[AR], seed=1405, run_id=2, stream_base=3,
                        precision=nat.F64, engine=nat.ENGINE_PERSISTENT)
    assert got["status"][0] == 0
    assert got["n_resampled"][0] == ref["n_resampled"]
    # north-star tolerance 1e-6 relative
    assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
    np.testing.assert_allclose(got["loglike_history"][0], ref["loglike_history"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-6)
    np.testing.assert_allclose(got["state_est"][0][:, 0], ref["state_est"][:, 0], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("ralg", [0, 1, 2])
def test_resample_algorithms_and_models(orc, engine, ralg):
    rng = np.random.default_rng(5)
    for model in (LG, RWD):
        y = sim_y(model, 20, rng)
        ref = orc.particle_filter(model, 0, ralg, 0, 3000, y, THETA[model], seed=9, stream=1)
        got = eh.filter_run(engine, model, 0, ralg, 0, 3000, y, THETA[model], seed=9, stream_base=1,
                            precision=nat.F64, engine=nat.ENGINE_PERSISTENT)
        assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
        np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-6)


def test_obs_times_and_early_exit(orc, engine):
    rng = np.random.default_rng(6)
    y = sim_y(AR, 6, rng)
    ot = [1, 2, 4, 7, 8, 12]
    ref = orc.particle_filter(AR, 0, 2, 0, 2048, y, THETA[AR], obs_times=ot, seed=3)
    got = eh.filter_run(engine, AR, 0, 2, 0, 2048, y, THETA[AR], obs_times=ot, seed=3, precision=nat.F64,
                        engine=nat.ENGINE_PERSISTENT)
    assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
    y2 = np.array([0.1, 1e6, 0.2])
    th = [0.8, 1.0, 1e-3]
    ref = orc.particle_filter(AR, 0, 2, 0, 512, y2, th, seed=3)
    got = eh.filter_run(engine, AR, 0, 2, 0, 512, y2, th, seed=3, precision=nat.F64, engine=nat.ENGINE_PERSISTENT)
    assert ref["early_exit"] == 1 and got["early_exit"][0] == 1 and got["loglike"][0] == -np.inf
    np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-6)


def test_batched_f32_against_kalman(orc, engine):
    # config C3 shape (reduced): linear-Gaussian, replicate filters, SISR; estimate within 3 MC standard errors
    rng = np.random.default_rng(7)
    y = sim_y(LG, 200, rng)
    exact = orc.kalman_loglik(y, 0.8, 1.0, 1.0)
    got = eh.filter_run(engine, LG, 0, 1, 0, 1 << 14, y, THETA[LG], seed=11, num_filters=64, precision=nat.F32,
                        engine=nat.ENGINE_PERSISTENT)
    assert (got["status"] == 0).all()
    lls = got["loglike"]
    est = np.log(np.mean(np.exp(lls - lls.max()))) + lls.max()
    se = lls.std(ddof=1) / np.sqrt(len(lls))
    assert abs(est - exact) < 3 * se + 0.01, (est, exact, se)
    assert len(np.unique(lls)) == 64


def test_f32_persistent_close_to_general_f32(engine):
    rng = np.random.default_rng(8)
    y = sim_y(AR, 50, rng)
    a = eh.filter_run(engine, AR, 0, 2, 0, 1 << 16, y, THETA[AR], seed=21, precision=nat.F32, engine=nat.ENGINE_PERSISTENT)
    b = eh.filter_run(engine, AR, 0, 2, 0, 1 << 16, y, THETA[AR], seed=21, precision=nat.F32, engine=nat.ENGINE_GENERAL)
    assert abs(a["loglike"][0] - b["loglike"][0]) < 0.25   # Monte-Carlo error at N = 2^16 over 50 steps
    np.testing.assert_allclose(a["state_est"][0], b["state_est"][0], atol=0.05)
    assert abs(int(a["n_resampled"][0]) - int(b["n_resampled"][0])) <= 1


def test_full_size_n_2pow20_runs_and_is_consistent(engine):
    # BASELINE config C2 geometry (whole chip, one filter), short T
    rng = np.random.default_rng(9)
    y = sim_y(AR, 20, rng)
    a = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F32, engine=nat.ENGINE_PERSISTENT)
    b = eh.filter_run(engine, AR, 0, 2, 1, 1 << 20, y, THETA[AR], seed=5, precision=nat.F32, engine=nat.ENGINE_PERSISTENT)
    g = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F32, engine=nat.ENGINE_GENERAL)
    # f32 runs are perturbed copies of one another (different rounding => a few different ancestors), so they
    # agree to Monte-Carlo error, which one low-ESS observation dominates (scripts/diag_precision.py): ~2e-2 here
    assert abs(a["loglike"][0] - g["loglike"][0]) < 0.1 and abs(b["loglike"][0] - g["loglike"][0]) < 0.1
    np.testing.assert_allclose(a["state_est"][0], g["state_est"][0], atol=0.02)
    # in f64 the persistent kernel and the general engine agree to rounding even at this size
    p64 = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F64, engine=nat.ENGINE_PERSISTENT)
    g64 = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F64, engine=nat.ENGINE_GENERAL)
    assert abs(p64["loglike"][0] - g64["loglike"][0]) <= 1e-9 * abs(g64["loglike"][0])
    np.testing.assert_allclose(p64["ess"][0], g64["ess"][0], rtol=1e-9)


def test_pmmh_on_persistent_engine_matches_oracle(orc, engine):
    from test_pmmh_gpu import PRIOR, readme_data
    rng = np.random.default_rng(1405)
    y = readme_data(12, rng)
    inits = np.array([[0.8, 1.0, 0.5], [0.5, 0.7, 1.2]])
    kw = dict(transform=[2, 1, 1], pilot_proposal_sd=[0.1, 0.15, 0.2], pilot_n=64, pilot_m=30, pilot_reps=6, m=40, seed=99)
    got = eh.pmmh_run(engine, 0, 0, y, inits, chain_id_base=4, engine=nat.ENGINE_PERSISTENT, **PRIOR, **kw)
    assert (got["status"] == 0).all()
    for c in range(2):
        ref = orc.pmmh_chain(0, 0, y, inits[c], chain_id=4 + c, **PRIOR, **kw)
        assert got["target_n"][c] == ref["target_n"]
        np.testing.assert_allclose(got["pilot_theta_chain"][c], ref["pilot_theta_chain"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got["theta_chain"][c], ref["theta_chain"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got["loglike_chain"][c], ref["loglike_chain"], rtol=1e-6)


def test_many_filters_per_group_and_several_groups_per_sm(orc, engine):
    # more filters than resident groups: every group runs several filters back to back (record buffers and x_new
    # are reused across filters); small slices put several groups on one SM
    rng = np.random.default_rng(12)
    y = sim_y(AR, 8, rng)
    C = 700
    got = eh.filter_run(engine, AR, 0, 2, 0, 2048, y, THETA[AR], seed=31, num_filters=C, precision=nat.F64,
                        engine=nat.ENGINE_PERSISTENT)
    assert (got["status"] == 0).all() and np.isfinite(got["loglike"]).all()
    for c in (0, 1, 147, 148, 333, 699):
        ref = orc.particle_filter(AR, 0, 2, 0, 2048, y, THETA[AR], seed=31, stream=c)
        assert abs(got["loglike"][c] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
        np.testing.assert_allclose(got["state_est"][c][:, 0], ref["state_est"][:, 0], rtol=1e-6, atol=1e-6)
    big = eh.filter_run(engine, AR, 0, 1, 1, 30000, y, THETA[AR], seed=32, num_filters=40, precision=nat.F32,
                        engine=nat.ENGINE_PERSISTENT)
    assert (big["status"] == 0).all() and np.isfinite(big["loglike"]).all() and (big["n_resampled"] == 8).all()


def test_auto_falls_back_to_the_general_kernels_when_a_fast_launch_is_refused(engine, monkeypatch):
    # the persistent kernel sizes its cooperative launch at launch time and may refuse (fewer resident CTAs than the choice
    # assumed); AUTO then runs the general kernels, a request by name stays an error
    y = sim_y(AR, 12, np.random.default_rng(9))
    ref = eh.filter_run(engine, AR, 0, 2, 0, 50000, y, THETA[AR], seed=3, precision=nat.F32, engine=nat.ENGINE_GENERAL)
    monkeypatch.setenv("BSSM_TEST_FAST_LAUNCH_UNSUPPORTED", "1")
    got = eh.filter_run(engine, AR, 0, 2, 0, 50000, y, THETA[AR], seed=3, precision=nat.F32, engine=nat.ENGINE_AUTO)
    assert got["status"][0] == 0 and got["loglike"][0] == ref["loglike"][0]
    np.testing.assert_array_equal(got["state_est"], ref["state_est"])
    with pytest.raises(Exception, match="refused"):
        eh.filter_run(engine, AR, 0, 2, 0, 50000, y, THETA[AR], seed=3, precision=nat.F32, engine=nat.ENGINE_PERSISTENT)
