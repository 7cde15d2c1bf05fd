"""The persistent bootstrap-filter kernel without a GPU: the kernel text of bayesssm_b200/csrc/bssm_fast.cuh (the whole
T loop in ONE cooperative launch; the CTAs of a group exchange epoch-tagged records and particles through global
memory and poll for them) is compiled by g++ over the SIMT emulation of tests/simt_emu.h -- every CTA on its own OS
thread, its threads as fibers, a poll yields -- and launched with the geometry fast_launch() computes
(tests/host_fast.cpp).  Compared with the oracle's Philox-mode filter (R/particle_filter_core.R:76-266 +
src/resampling.cpp:16-66 restated in oracle/pf_oracle.c): 1e-9 in the parity precision (the GPU tests allow 1e-6;
only summation order and FMA contraction differ), statistical closeness in the throughput precision (libm here
where the device uses the SFU)."""
import os
import subprocess

import numpy as np
import pytest

from test_filter_gpu import THETA, sim_y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AR, LG, RWD, ARCOS = 0, 1, 2, 4
F64_HEADS, F64_LOOPS, F32_HEADS, F32_LOOPS = 0, 1, 2, 3     # the four instantiations fast_model() chooses from


@pytest.fixture(scope="module")
def host_fast(tmp_path_factory):
    exe = tmp_path_factory.mktemp("hf") / "host_fast"
    subprocess.run(["g++", "-O1", "-std=c++20", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_fast.cpp")], check=True)

    def run(model, variant, G, N, y, thetas, ngroups=1, resample_fn=0, resample_algorithm=2, threshold=-1.0, seed=1405,
            run_id=2, stream_base=3, n_per=(), obs_times=None):
        y = np.ascontiguousarray(y, dtype=np.float64)
        th = np.zeros((len(thetas), 3))
        for c, t in enumerate(thetas):
            th[c, :len(t)] = t
        args = [model, variant, G, ngroups, N, len(y), len(thetas), resample_fn, resample_algorithm, threshold, seed, run_id, stream_base] + list(n_per)
        r = subprocess.run([str(exe)] + [str(a) for a in args], input=y.tobytes() + th.tobytes(), capture_output=True, timeout=600,
                           env=dict(os.environ, EMU_OBS_TIMES=",".join(str(int(t)) for t in obs_times)) if obs_times is not None else None)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        lines, recs = r.stdout.decode().strip().splitlines(), []
        for i in range(0, len(lines), 4):
            h = lines[i].split()
            recs.append({"filter": int(h[3]), "loglike": float(h[5]), "n_resampled": int(h[7]), "status": int(h[9]),
                         "early_exit": int(h[11]), "ess": np.array(lines[i + 1].split()[1:], float),
                         "state_est": np.array(lines[i + 2].split()[1:], float),
                         "loglike_history": np.array(lines[i + 3].split()[1:], float)})
        return recs
    return run


def check(rec, ref, tol=1e-9):
    assert rec["status"] == 0 and rec["early_exit"] == ref["early_exit"]
    assert rec["n_resampled"] == ref["n_resampled"]
    assert abs(rec["loglike"] - ref["loglike"]) <= tol * abs(ref["loglike"])
    np.testing.assert_allclose(rec["loglike_history"], ref["loglike_history"], rtol=tol, atol=tol)
    np.testing.assert_allclose(rec["ess"], ref["ess"], rtol=tol)
    np.testing.assert_allclose(rec["state_est"], ref["state_est"][:, 0], rtol=tol, atol=tol)


# group sizes 1 - 4, slices that are ragged / smaller than a warp / exact, both expansions, both resamplers
@pytest.mark.parametrize("variant,G,N,T,rfn", [(F64_HEADS, 1, 1, 5, 0), (F64_HEADS, 2, 3, 6, 1), (F64_HEADS, 1, 1000, 6, 0),
                                               (F64_HEADS, 4, 1025, 6, 1), (F64_HEADS, 3, 3000, 8, 0), (F64_LOOPS, 2, 5000, 5, 0)])
def test_f64_kernel_text_matches_oracle(orc, host_fast, variant, G, N, T, rfn):
    y = sim_y(AR, T, np.random.default_rng(N))
    ref = orc.particle_filter(AR, 0, 2, rfn, N, y, THETA[AR], seed=1405, run_id=2, stream=3)
    rec, = host_fast(AR, variant, G, N, y, [THETA[AR]], resample_fn=rfn)
    check(rec, ref)


@pytest.mark.parametrize("model,variant,ralg", [(LG, F64_HEADS, 2), (RWD, F64_HEADS, 0), (ARCOS, F64_LOOPS, 1), (AR, F64_LOOPS, 1)])
def test_models_resample_algorithms_and_groups_walking_several_filters(orc, host_fast, model, variant, ralg):
    y = sim_y(AR if model == ARCOS else model, 6, np.random.default_rng(5))
    base = np.array(THETA[AR] if model == ARCOS else THETA[model])
    thetas = [list(base * (1 + 0.05 * c)) for c in range(3)]
    recs = host_fast(model, variant, 3, 2000, y, thetas, ngroups=2, resample_algorithm=ralg, seed=9, run_id=0, stream_base=1)
    for c, rec in enumerate(recs):     # two groups share three filters: filter c keeps its own theta and Philox stream
        check(rec, orc.particle_filter(model, 0, ralg, 0, 2000, y, thetas[c], seed=9, stream=1 + c))


def test_ragged_batch_particle_counts_per_filter(orc, host_fast):
    # PMMH's tuned target_n differs from chain to chain (R/pmmh_tuning.R:54-57): FilterDev::n_per
    y = sim_y(AR, 6, np.random.default_rng(3))
    ns = [3000, 50, 1777]
    thetas = [list(np.array(THETA[AR]) * (1 + 0.05 * c)) for c in range(3)]
    recs = host_fast(AR, F64_HEADS, 3, 3000, y, thetas, ngroups=2, seed=5, run_id=1, stream_base=2, n_per=ns)
    for c, rec in enumerate(recs):
        check(rec, orc.particle_filter(AR, 0, 2, 0, ns[c], y, thetas[c], seed=5, run_id=1, stream=2 + c))


def test_threshold_early_exit_and_no_observations(orc, host_fast):
    y = sim_y(AR, 6, np.random.default_rng(6))
    rec, = host_fast(AR, F64_HEADS, 2, 2048, y, [THETA[AR]], threshold=1500.0, seed=3, run_id=0, stream_base=0)
    check(rec, orc.particle_filter(AR, 0, 2, 0, 2048, y, THETA[AR], seed=3, threshold=1500.0))
    y2, th = np.array([0.1, 1e6, 0.2]), [0.8, 1.0, 1e-3]               # R/particle_filter_core.R:189-202
    rec, = host_fast(AR, F64_HEADS, 2, 512, y2, [th], seed=3, run_id=0, stream_base=0)
    ref = orc.particle_filter(AR, 0, 2, 0, 512, y2, th, seed=3)
    assert ref["early_exit"] == 1 and rec["early_exit"] == 1 and rec["loglike"] == -np.inf and rec["status"] == 0
    np.testing.assert_allclose(rec["ess"], ref["ess"], rtol=1e-9)
    rec, = host_fast(AR, F64_HEADS, 2, 777, np.zeros(0), [THETA[AR]], seed=3, run_id=0, stream_base=0)
    ref = orc.particle_filter(AR, 0, 2, 0, 777, np.zeros(0), THETA[AR], seed=3)
    assert rec["ess"][0] == 777 and abs(rec["state_est"][0] - ref["state_est"][0, 0]) < 1e-12


@pytest.mark.parametrize("variant,G,N", [(F64_HEADS, 4, 6000), (F64_LOOPS, 3, 20000)])
def test_degenerate_weights_a_few_particles_take_everything(orc, host_fast, variant, G, N):
    # a very sharp likelihood: offspring of one source spread over the slices of several CTAs, several expansion passes
    y, th = sim_y(AR, 5, np.random.default_rng(11)), [0.8, 1.0, 2e-4]
    rec, = host_fast(AR, variant, G, N, y, [th], resample_algorithm=1, seed=13, run_id=0, stream_base=0)
    check(rec, orc.particle_filter(AR, 0, 1, 0, N, y, th, seed=13))


def test_result_does_not_depend_on_the_group_size(host_fast):
    y = sim_y(AR, 6, np.random.default_rng(2))
    a, = host_fast(AR, F64_HEADS, 2, 4096, y, [THETA[AR]])
    b, = host_fast(AR, F64_HEADS, 5, 4096, y, [THETA[AR]])
    assert a["n_resampled"] == b["n_resampled"]
    np.testing.assert_allclose(a["loglike_history"], b["loglike_history"], rtol=1e-12)
    np.testing.assert_allclose(a["state_est"], b["state_est"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("variant,G,N", [(F32_HEADS, 3, 6000), (F32_LOOPS, 3, 20000)])
def test_f32_kernel_text_is_close(orc, host_fast, variant, G, N):
    y = sim_y(AR, 5, np.random.default_rng(N))
    ref = orc.particle_filter(AR, 0, 2, 0, N, y, THETA[AR], seed=1405, run_id=2, stream=3)
    rec, = host_fast(AR, variant, G, N, y, [THETA[AR]])
    assert rec["status"] == 0 and rec["n_resampled"] == ref["n_resampled"]
    assert abs(rec["loglike"] - ref["loglike"]) < 2e-2 and np.abs(rec["state_est"] - ref["state_est"][:, 0]).max() < 2e-2


def test_nan_observation_is_reported_as_in_the_oracle(orc, host_fast):
    # NaN weights: R's `if (all(lw < -1e8))` raises "missing value where TRUE/FALSE needed" (R/particle_filter_core.R:189);
    # here status 3 (BSSM_ERR_NAN_WEIGHT), the log-likelihood staying at the last finite observation
    y = sim_y(AR, 5, np.random.default_rng(1))
    y[2] = np.nan
    ref = orc.particle_filter(AR, 0, 2, 0, 3000, y, THETA[AR], seed=1)
    assert ref["status"] == 3
    for rec in host_fast(AR, F64_HEADS, 3, 3000, y, [THETA[AR]], seed=1, run_id=0, stream_base=0):
        assert rec["status"] == 3 and rec["loglike"] == pytest.approx(ref["loglike"], rel=1e-12)


def test_observation_times_with_gaps(orc, host_fast):
    # R/particle_filter_core.R:70-71,124-136: the gap to the previous observation time is that many transitions
    y, ot = sim_y(AR, 6, np.random.default_rng(6)), [1, 2, 4, 7, 8, 12]
    ref = orc.particle_filter(AR, 0, 2, 0, 2048, y, THETA[AR], obs_times=ot, seed=3)
    rec, = host_fast(AR, F64_HEADS, 3, 2048, y, [THETA[AR]], seed=3, run_id=0, stream_base=0, obs_times=ot)
    check(rec, ref)


def test_outlying_observations_redo_the_step_against_the_true_maximum(orc, host_fast):
    # observations 40 - 60 standard deviations away from every particle: log-weights around -1000 to -2000, far from the early-exit
    # threshold (-1e8) but deep in exp's underflow range unless the maximum is taken out first
    y = np.array([0.1, 30.0, 0.2, -25.0, 0.3, 0.1])
    ref = orc.particle_filter(AR, 0, 2, 0, 3000, y, THETA[AR], seed=7)
    rec, = host_fast(AR, F64_HEADS, 3, 3000, y, [THETA[AR]], seed=7, run_id=0, stream_base=0)
    check(rec, ref)
    rec, = host_fast(AR, F32_LOOPS, 2, 9000, y, [THETA[AR]], seed=7, run_id=0, stream_base=0)
    ref = orc.particle_filter(AR, 0, 2, 0, 9000, y, THETA[AR], seed=7)
    assert rec["status"] == 0 and rec["n_resampled"] == ref["n_resampled"]
    np.testing.assert_allclose(rec["loglike_history"], ref["loglike_history"], rtol=2e-4)


def test_more_than_32_ctas_in_a_group(orc, host_fast):
    # the records are scanned in groups of 32 CTAs; 40 CTAs make two groups
    y = sim_y(AR, 5, np.random.default_rng(40))
    ref = orc.particle_filter(AR, 0, 2, 0, 2560, y, THETA[AR], seed=4)
    rec, = host_fast(AR, F64_HEADS, 40, 2560, y, [THETA[AR]], seed=4, run_id=0, stream_base=0)
    check(rec, ref)

