"""The streaming bootstrap-filter engine (bssm_stream.cuh: particles in HBM, two kernels per observation)
against the oracle and the other engines.  Same Philox streams on every side, so in f64 it reproduces the
oracle's Philox-mode filter up to summation order; in f32 (throughput precision) it is checked
statistically (Kalman) and against the general engine."""
import numpy as np
import pytest

import engine_helpers as eh
from bayesssm_b200 import _native as nat
from test_filter_gpu import THETA, sim_y

pytestmark = pytest.mark.gpu
AR, LG, RWD, ARCOS = 0, 1, 2, 4
ST = nat.ENGINE_STREAM


# sizes around the tile boundaries (f64 tile = 1024 particles): one partial tile, exact tiles, ragged tails
@pytest.mark.parametrize("N,T", [(1, 5), (3, 6), (1000, 30), (1024, 10), (1025, 10), (4096, 40), (70001, 12), (1 << 17, 6)])
@pytest.mark.parametrize("rfn", [0, 1])
def test_f64_matches_oracle_philox(orc, engine, N, T, rfn):
    rng = np.random.default_rng(N + rfn)
    y = sim_y(AR, T, rng)
    ref = orc.particle_filter(AR, 0, 2, rfn, N, y, THETA[AR], seed=1405, run_id=2, stream=3)
    got = eh.filter_run(engine, AR, 0, 2, rfn, N, y, THETA[AR], seed=1405, run_id=2, stream_base=3,
                        precision=nat.F64, engine=ST)
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
    for model in (LG, RWD, ARCOS):
        y = sim_y(AR if model == ARCOS else model, 20, rng)
        th = THETA[AR] if model == ARCOS else THETA[model]
        ref = orc.particle_filter(model, 0, ralg, 0, 3000, y, th, seed=9, stream=1)
        got = eh.filter_run(engine, model, 0, ralg, 0, 3000, y, th, seed=9, stream_base=1, precision=nat.F64, engine=ST)
        assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
        np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-6)
        np.testing.assert_allclose(got["state_est"][0][:, 0], ref["state_est"][:, 0], rtol=1e-6, atol=1e-6)


def test_threshold_obs_times_and_early_exit(orc, engine):
    rng = np.random.default_rng(6)
    y = sim_y(AR, 6, rng)
    ot = [1, 2, 4, 7, 8, 12]
    ref = orc.particle_filter(AR, 0, 2, 0, 2048, y, THETA[AR], obs_times=ot, seed=3, threshold=1500.0)
    got = eh.filter_run(engine, AR, 0, 2, 0, 2048, y, THETA[AR], obs_times=ot, seed=3, threshold=1500.0,
                        precision=nat.F64, engine=ST)
    assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
    assert got["n_resampled"][0] == ref["n_resampled"]
    y2 = np.array([0.1, 1e6, 0.2])
    th = [0.8, 1.0, 1e-3]
    ref = orc.particle_filter(AR, 0, 2, 0, 512, y2, th, seed=3)
    got = eh.filter_run(engine, AR, 0, 2, 0, 512, y2, th, seed=3, precision=nat.F64, engine=ST)
    assert ref["early_exit"] == 1 and got["early_exit"][0] == 1 and got["loglike"][0] == -np.inf
    np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-6)
    # T = 0: only the initial state estimate
    got = eh.filter_run(engine, AR, 0, 2, 0, 777, np.zeros((0, 1)), THETA[AR], seed=3, precision=nat.F64, engine=ST)
    ref = orc.particle_filter(AR, 0, 2, 0, 777, np.zeros(0), THETA[AR], seed=3)
    assert got["ess"][0][0] == 777 and abs(got["state_est"][0][0, 0] - ref["state_est"][0, 0]) < 1e-12


def test_degenerate_weights_one_particle_takes_everything(orc, engine):
    # a very sharp likelihood: a handful of particles own all offspring (heavy-source path, many output chunks
    # from one tile, empty tiles elsewhere)
    rng = np.random.default_rng(11)
    y = sim_y(AR, 5, rng)
    th = [0.8, 1.0, 2e-4]
    for prec, tol in ((nat.F64, 1e-6), (nat.F32, None)):
        ref = orc.particle_filter(AR, 0, 1, 0, 50000, y, th, seed=13)
        got = eh.filter_run(engine, AR, 0, 1, 0, 50000, y, th, seed=13, precision=prec, engine=ST)
        assert got["status"][0] == 0 and got["early_exit"][0] == ref["early_exit"]
        if tol:
            assert abs(got["loglike"][0] - ref["loglike"]) <= tol * abs(ref["loglike"])
            np.testing.assert_allclose(got["state_est"][0][:, 0], ref["state_est"][:, 0], rtol=1e-6, atol=1e-6)
        else:
            assert abs(got["loglike"][0] - ref["loglike"]) <= 0.1   # f32 at sigma_y = 2e-4: weights themselves carry ~1e-3 errors
            np.testing.assert_allclose(got["state_est"][0][:, 0], ref["state_est"][:, 0], atol=2e-3)


def test_batched_f32_against_kalman(orc, engine):
    # config C3 shape (reduced): linear-Gaussian, replicate filters, SISR; estimate within 3 MC standard errors
    rng = np.random.default_rng(7)
    y = sim_y(LG, 200, rng)
    exact = orc.kalman_loglik(y, 0.8, 1.0, 1.0)
    got = eh.filter_run(engine, LG, 0, 1, 0, 1 << 14, y, THETA[LG], seed=11, num_filters=64, precision=nat.F32, engine=ST)
    assert (got["status"] == 0).all()
    lls = got["loglike"]
    est = np.log(np.mean(np.exp(lls - lls.max()))) + lls.max()
    se = lls.std(ddof=1) / np.sqrt(len(lls))
    assert abs(est - exact) < 3 * se + 0.01, (est, exact, se)
    assert len(np.unique(lls)) == 64


def test_batched_filters_match_oracle_per_stream(orc, engine):
    rng = np.random.default_rng(12)
    y = sim_y(AR, 8, rng)
    C = 300
    got = eh.filter_run(engine, AR, 0, 2, 0, 2500, y, THETA[AR], seed=31, num_filters=C, precision=nat.F64, engine=ST)
    assert (got["status"] == 0).all() and np.isfinite(got["loglike"]).all()
    for c in (0, 1, 147, 299):
        ref = orc.particle_filter(AR, 0, 2, 0, 2500, y, THETA[AR], seed=31, stream=c)
        assert abs(got["loglike"][c] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
        np.testing.assert_allclose(got["state_est"][c][:, 0], ref["state_est"][:, 0], rtol=1e-6, atol=1e-6)


def test_f32_close_to_general_and_f64_identical_at_2pow20(engine):
    # BASELINE config C2 geometry, short T
    rng = np.random.default_rng(9)
    y = sim_y(AR, 20, rng)
    a = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F32, engine=ST)
    b = eh.filter_run(engine, AR, 0, 2, 1, 1 << 20, y, THETA[AR], seed=5, precision=nat.F32, engine=ST)
    g = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F32, engine=nat.ENGINE_GENERAL)
    assert abs(a["loglike"][0] - g["loglike"][0]) < 0.1 and abs(b["loglike"][0] - g["loglike"][0]) < 0.1
    np.testing.assert_allclose(a["state_est"][0], g["state_est"][0], atol=0.02)
    s64 = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F64, engine=ST)
    g64 = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F64, engine=nat.ENGINE_GENERAL)
    assert abs(s64["loglike"][0] - g64["loglike"][0]) <= 1e-9 * abs(g64["loglike"][0])
    np.testing.assert_allclose(s64["ess"][0], g64["ess"][0], rtol=1e-9)
    np.testing.assert_allclose(s64["state_est"][0], g64["state_est"][0], rtol=1e-9, atol=1e-9)


def test_larger_than_the_persistent_kernel_holds(engine):
    # N = 2^23 exceeds the register-resident kernel (<= 256 x 7168); AUTO must pick the streaming engine
    rng = np.random.default_rng(10)
    y = sim_y(AR, 6, rng)
    n0 = engine.launch_count()
    a = eh.filter_run(engine, AR, 0, 2, 0, 1 << 23, y, THETA[AR], seed=5, precision=nat.F32, engine=nat.ENGINE_AUTO)
    assert engine.launch_count() - n0 < 40          # 2 kernels per observation, not the general engine's ~10
    s = eh.filter_run(engine, AR, 0, 2, 0, 1 << 20, y, THETA[AR], seed=5, precision=nat.F32, engine=ST)
    assert a["status"][0] == 0 and abs(a["loglike"][0] - s["loglike"][0]) < 0.05


def test_pmmh_on_streaming_engine_matches_oracle(orc, engine):
    from test_pmmh_gpu import PRIOR, readme_data
    rng = np.random.default_rng(1405)
    y = readme_data(12, rng)
    inits = np.array([[0.8, 1.0, 0.5], [0.5, 0.7, 1.2]])
    kw = dict(transform=[2, 1, 1], pilot_proposal_sd=[0.1, 0.15, 0.2], pilot_n=64, pilot_m=30, pilot_reps=6, m=40, seed=99)
    got = eh.pmmh_run(engine, 0, 0, y, inits, chain_id_base=4, engine=ST, **PRIOR, **kw)
    assert (got["status"] == 0).all()
    for c in range(2):
        ref = orc.pmmh_chain(0, 0, y, inits[c], chain_id=4 + c, **PRIOR, **kw)
        assert got["target_n"][c] == ref["target_n"]
        np.testing.assert_allclose(got["pilot_theta_chain"][c], ref["pilot_theta_chain"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got["theta_chain"][c], ref["theta_chain"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(got["loglike_chain"][c], ref["loglike_chain"], rtol=1e-6)


@pytest.mark.parametrize("eng", [nat.ENGINE_GENERAL, nat.ENGINE_PERSISTENT, nat.ENGINE_STREAM])
def test_nan_weights_are_reported_by_every_engine(engine, eng):
    # R's `if (all(lw < -1e8))` raises "missing value where TRUE/FALSE needed" on NaN weights (R/particle_filter_core.R:189):
    # status BSSM_ERR_NAN_WEIGHT, whichever engine runs -- also when a single particle of a batch member is affected
    rng = np.random.default_rng(3)
    y = sim_y(AR, 5, rng)
    got = eh.filter_run(engine, AR, 0, 2, 0, 5000, y, [0.8, 1.0, float("nan")], seed=1, precision=nat.F32, engine=eng)
    assert got["status"][0] == nat.ERR_NAN_WEIGHT
    th = np.tile(np.array([0.8, 1.0, 0.5]), (3, 1))
    th[1, 0] = float("nan")                  # NaN dynamics: the states, hence the weights, of filter 1 only
    got = eh.filter_run(engine, AR, 0, 2, 0, 5000, y, th, seed=1, num_filters=3, precision=nat.F32, engine=eng)
    assert got["status"].tolist() == [0, nat.ERR_NAN_WEIGHT, 0]
    assert np.isfinite(got["loglike"][[0, 2]]).all()


def test_runs_are_bit_reproducible(engine):
    # no atomics on floating-point data, fixed merge order, every output slot written exactly once: repeated runs are
    # bit-identical (a race between tiles / blocks would show up here), for one big filter and for a batch
    rng = np.random.default_rng(4)
    y = sim_y(AR, 30, rng)
    for kw in (dict(N=1 << 21, num_filters=1), dict(N=40000, num_filters=37)):
        runs = [eh.filter_run(engine, AR, 0, 2, 0, kw["N"], y, THETA[AR], seed=17, num_filters=kw["num_filters"], precision=nat.F32,
                              engine=ST) for _ in range(3)]
        for r in runs[1:]:
            assert np.array_equal(r["loglike"], runs[0]["loglike"])
            assert np.array_equal(r["state_est"], runs[0]["state_est"]) and np.array_equal(r["ess"], runs[0]["ess"])
        assert (runs[0]["status"] == 0).all() and (runs[0]["n_resampled"] > 0).all()


def test_big_single_filters_f32_against_kalman(orc, engine):
    # one filter of 2^22 particles (256-thread variant, the size class of the particle-sharded runs), throughput
    # precision, SISR (under SISAR the reference drops the weights of non-resampled steps -- quirk A1 -- and the
    # estimate is biased by construction): the estimates of a few independent runs agree with the exact Kalman value
    rng = np.random.default_rng(21)
    y = sim_y(LG, 100, rng)
    exact = orc.kalman_loglik(y, 0.8, 1.0, 1.0)
    lls = np.array([eh.filter_run(engine, LG, 0, 1, 0, 1 << 22, y, THETA[LG], seed=100 + s, precision=nat.F32, engine=ST)["loglike"][0]
                    for s in range(6)])
    se = lls.std(ddof=1) / np.sqrt(len(lls))
    assert abs(lls.mean() - exact) < 3 * se + 5e-3, (lls, exact)
    assert lls.std(ddof=1) < 0.05


# multinomial resampling on the streaming engine (sorted uniforms from exponential spacings; oracle resample_fn = 3)
@pytest.mark.parametrize("N,T,ralg", [(1000, 20, 2), (4097, 12, 1), (70001, 10, 2), (1 << 17, 6, 1)])
def test_multinomial_f64_matches_oracle(orc, engine, N, T, ralg):
    y = sim_y(AR, T, np.random.default_rng(N))
    ref = orc.particle_filter(AR, 0, ralg, 3, N, y, THETA[AR], seed=1405, run_id=2, stream=3)
    got = eh.filter_run(engine, AR, 0, ralg, 2, N, y, THETA[AR], seed=1405, run_id=2, stream_base=3,
                        precision=nat.F64, engine=nat.ENGINE_STREAM)
    assert got["status"][0] == 0 and got["n_resampled"][0] == ref["n_resampled"]
    assert abs(got["loglike"][0] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])      # north-star tolerance
    np.testing.assert_allclose(got["ess"][0], ref["ess"], rtol=1e-6)
    np.testing.assert_allclose(got["state_est"][0][:, 0], ref["state_est"][:, 0], rtol=1e-6, atol=1e-6)


def test_multinomial_f32_batched_against_kalman_and_auto_engine(orc, engine):
    y = sim_y(LG, 200, np.random.default_rng(7))
    exact = orc.kalman_loglik(y, 0.8, 1.0, 1.0)
    got = eh.filter_run(engine, LG, 0, 1, 2, 1 << 14, y, THETA[LG], seed=11, num_filters=64, precision=nat.F32, engine=nat.ENGINE_AUTO)
    assert (got["status"] == 0).all()
    lls = got["loglike"]
    est = np.log(np.mean(np.exp(lls - lls.max()))) + lls.max()
    se = lls.std(ddof=1) / np.sqrt(len(lls))
    assert abs(est - exact) < 3 * se + 0.01, (est, exact, se)
    # AUTO took the streaming engine: the same numbers when it is asked for by name
    again = eh.filter_run(engine, LG, 0, 1, 2, 1 << 14, y, THETA[LG], seed=11, num_filters=64, precision=nat.F32, engine=nat.ENGINE_STREAM)
    np.testing.assert_array_equal(again["loglike"], lls)


# ---- batches: the chain-persistent kernel (k_st_chain, one cooperative launch per group of filters) ----
@pytest.mark.parametrize("C,N,T,prec", [(24, 70001, 12, nat.F64), (1500, 3000, 10, nat.F32), (128, 65536, 40, nat.F32), (17, 1, 5, nat.F64)])
def test_chain_persistent_kernel_equals_the_launch_per_body_form(engine, monkeypatch, C, N, T, prec):
    # same bodies, same block -> tile ranges when the blocks per filter agree: bit-identical results, several launch groups
    # (1500 filters > resident block slots), ragged tiles, SISAR decisions taken per filter
    y = sim_y(AR, T, np.random.default_rng(C))
    out = {}
    for chain in ("1", "0"):
        monkeypatch.setenv("BSSM_ST_CHAIN", chain)
        monkeypatch.setenv("BSSM_ST_BPC", "3" if N > 3000 else "1")
        out[chain] = eh.filter_run(engine, AR, 0, 2, 0, N, y, THETA[AR], seed=77, num_filters=C, precision=prec, engine=ST)
    a, b = out["1"], out["0"]
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    np.testing.assert_array_equal(a["loglike"], b["loglike"])
    np.testing.assert_array_equal(a["n_resampled"], b["n_resampled"])
    np.testing.assert_array_equal(a["ess"], b["ess"])
    np.testing.assert_array_equal(a["state_est"], b["state_est"])
    assert len(np.unique(a["loglike"])) == C


def test_chain_persistent_kernel_against_oracle_with_dead_filters_and_gaps(orc, engine, monkeypatch):
    monkeypatch.setenv("BSSM_ST_CHAIN", "1")
    y = sim_y(AR, 9, np.random.default_rng(2))
    times = [1, 2, 4, 5, 6, 9, 10, 11, 12]
    got = eh.filter_run(engine, AR, 0, 2, 1, 5000, y, THETA[AR], seed=13, num_filters=40, precision=nat.F64, engine=ST, obs_times=times)
    for c in (0, 7, 39):
        ref = orc.particle_filter(AR, 0, 2, 1, 5000, y, THETA[AR], seed=13, stream=c, obs_times=times)
        assert abs(got["loglike"][c] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"]) and got["n_resampled"][c] == ref["n_resampled"]
        np.testing.assert_allclose(got["ess"][c], ref["ess"], rtol=1e-6)
    # all weights below -1e8 at the second observation: every filter stops there (R/particle_filter_core.R:189-202)
    y2, th = np.array([0.1, 1e6, 0.2]), [0.8, 1.0, 1e-3]
    got = eh.filter_run(engine, AR, 0, 2, 0, 4096, y2, th, seed=3, num_filters=20, precision=nat.F64, engine=ST)
    assert (got["early_exit"] == 1).all() and (got["loglike"] == -np.inf).all() and (got["status"] == 0).all()


def test_pmmh_with_ragged_tuned_particle_counts_on_the_chain_kernel(orc, engine, monkeypatch):
    # 20 chains: the batched filters of the pilot chain, of .pilot_run's replicates (chains x reps filters) and of the main chain
    # -- where every chain has its own tuned particle count (FilterDev::n_per, R/pmmh_tuning.R:54-57) -- all run on k_st_chain;
    # the draws equal those of the launch-per-body form bit for bit, and two of the chains equal the oracle's
    from test_pmmh_gpu import PRIOR, readme_data
    rng = np.random.default_rng(77)
    y = readme_data(10, rng)
    inits = np.array([[0.8, 1.0, 0.5]] * 20) * (1 + 0.005 * np.arange(20))[:, None]
    kw = dict(transform=[2, 1, 1], pilot_proposal_sd=[0.1, 0.15, 0.2], pilot_n=64, pilot_m=20, pilot_reps=4, m=25, seed=5)
    out = {}
    for chain in ("1", "0"):
        monkeypatch.setenv("BSSM_ST_CHAIN", chain)
        out[chain] = eh.pmmh_run(engine, 0, 0, y, inits, chain_id_base=2, engine=ST, **PRIOR, **kw)
        assert (out[chain]["status"] == 0).all()
    a, b = out["1"], out["0"]
    assert len(set(a["target_n"].tolist())) > 1                      # ragged particle counts in the main phase
    np.testing.assert_array_equal(a["target_n"], b["target_n"])
    np.testing.assert_array_equal(a["theta_chain"], b["theta_chain"])
    np.testing.assert_array_equal(a["loglike_chain"], b["loglike_chain"])
    for c in (0, 19):
        ref = orc.pmmh_chain(0, 0, y, inits[c], chain_id=2 + c, **PRIOR, **kw)
        assert a["target_n"][c] == ref["target_n"]
        np.testing.assert_allclose(a["theta_chain"][c], ref["theta_chain"], rtol=1e-6, atol=1e-9)
