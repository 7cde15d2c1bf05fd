"""The general engine without a GPU: the kernel text of bayesssm_b200/csrc/bssm_filter.cuh (k_init, k_weight,
k_finalize, k_search_gather, k_flip, k_post) and bssm_resample.cuh (k_tile_sums / k_tile_scan / k_chain / k_tile_exact,
i.e. the bit-exact sequential cumsum of src/resampling.cpp:25 done in parallel) compiled by g++ over the SIMT emulation
(tests/simt_emu.h) and driven as run_filter_steps() / resample_stage() / resample_cdf() in bssm_engine.cu drive it
(tests/host_general.cpp).  Bootstrap, auxiliary (R/particle_filter_core.R:140-175) and resample-move (:226-234)
filters, every built-in model, the three resamplers; compared with the oracle's Philox-mode filter at 1e-9."""
import os
import subprocess

import numpy as np
import pytest

from test_filter_gpu import THETA, sim_y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AR, LG, RWD, SIR, ARCOS, RW2D, GIL = 0, 1, 2, 3, 4, 5, 6
NTHETA = {AR: 3, LG: 3, RWD: 2, SIR: 2, ARCOS: 3, RW2D: 1, GIL: 2}
BPF, APF, RMPF = 0, 1, 2


@pytest.fixture(scope="module")
def host_general(tmp_path_factory):
    exe = tmp_path_factory.mktemp("hg") / "host_general"
    subprocess.run(["g++", "-O1", "-std=c++20", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_general.cpp")], check=True)

    def run(model, algorithm, N, y, thetas, resample_fn=0, resample_algorithm=2, threshold=-1.0, seed=77, run_id=1, stream_base=5,
            exact=1, obs_times=None, prefix_tiles=None, carry=False):
        y = np.ascontiguousarray(y, dtype=np.float64)
        args = [model, algorithm, N, len(y), len(thetas), resample_fn, resample_algorithm, threshold, seed, run_id, stream_base, exact]
        env = dict(os.environ)
        if obs_times is not None:
            env["EMU_OBS_TIMES"] = ",".join(str(int(t)) for t in obs_times)
        if carry:
            env["EMU_CARRY"] = "1"
        if prefix_tiles is not None:       # tile totals scanned by k_tile_prefix from this many tiles on (large-input path of the cdf pipeline)
            env["EMU_RS_PREFIX_TILES"] = str(prefix_tiles)
        r = subprocess.run([str(exe)] + [str(a) for a in args], input=y.tobytes() + np.ascontiguousarray(thetas, dtype=np.float64).tobytes(),
                           capture_output=True, timeout=600, env=env)
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
    np.testing.assert_allclose(rec["state_est"], ref["state_est"].ravel(), rtol=tol, atol=tol)


@pytest.mark.parametrize("algorithm", [BPF, APF, RMPF])
@pytest.mark.parametrize("rfn", [0, 1, 2])
def test_algorithms_and_resamplers_on_the_readme_model(orc, host_general, algorithm, rfn):
    y = sim_y(AR, 5, np.random.default_rng(1))
    ref = orc.particle_filter(AR, algorithm, 2, rfn, 2500, y, THETA[AR], seed=1405, run_id=2, stream=3)
    rec, = host_general(AR, algorithm, 2500, y, [THETA[AR]], resample_fn=rfn, seed=1405, run_id=2, stream_base=3)
    check(rec, ref)


@pytest.mark.parametrize("model,algorithm,N,C,ralg", [(SIR, BPF, 2000, 1, 2), (SIR, APF, 1500, 1, 2), (SIR, RMPF, 1500, 1, 1),
                                                      (RW2D, BPF, 2500, 2, 2), (LG, APF, 1025, 2, 2), (RWD, RMPF, 700, 1, 0),
                                                      (ARCOS, BPF, 1, 1, 2), (GIL, BPF, 600, 2, 2), (GIL, APF, 500, 1, 2), (GIL, RMPF, 500, 1, 1)])
def test_every_builtin_model(orc, host_general, model, algorithm, N, C, ralg):
    # integer-valued two-dimensional SIR states with constants, two-dimensional random walk, batches with their own theta
    # GIL: the same SIR model with the exact (Gillespie) daily step, uniforms drawn on demand (DynU)
    base = THETA[SIR] if model == GIL else THETA[model]
    y = sim_y(SIR if model == GIL else model, 5, np.random.default_rng(model * 10 + algorithm))
    nth = NTHETA[model]
    thetas = [list(np.array(base[:nth]) * (1 + 0.05 * c)) + list(base[nth:]) for c in range(C)]
    recs = host_general(model, algorithm, N, y, thetas, resample_algorithm=ralg)
    for c, rec in enumerate(recs):
        th = thetas[c] if model in (SIR, GIL) else thetas[c][:nth]
        check(rec, orc.particle_filter(model, algorithm, ralg, 0, N, y, th, seed=77, run_id=1, stream=5 + c))


def test_plain_parallel_scan_path(orc, host_general):
    # exact_resampling = 0 (the throughput mode's cdf: ordinary fp64 scan): the same filter up to cdf rounding
    y = sim_y(AR, 5, np.random.default_rng(4))
    ref = orc.particle_filter(AR, BPF, 1, 0, 3000, y, THETA[AR], seed=77, run_id=1, stream=5)
    rec, = host_general(AR, BPF, 3000, y, [THETA[AR]], resample_algorithm=1, exact=0)
    assert rec["status"] == 0 and rec["n_resampled"] == ref["n_resampled"]
    assert abs(rec["loglike"] - ref["loglike"]) < 1e-3 and np.abs(rec["state_est"] - ref["state_est"].ravel()).max() < 1e-2


def test_nan_observation_is_reported_as_in_the_oracle(orc, host_general):
    # NaN weights: R's `if (all(lw < -1e8))` raises "missing value where TRUE/FALSE needed" (R/particle_filter_core.R:189);
    # here status 3 (BSSM_ERR_NAN_WEIGHT), the log-likelihood staying at the last finite observation
    y = sim_y(AR, 5, np.random.default_rng(1))
    y[2] = np.nan
    ref = orc.particle_filter(AR, 0, 2, 0, 3000, y, THETA[AR], seed=1)
    assert ref["status"] == 3
    for rec in host_general(AR, BPF, 3000, y, [THETA[AR]], seed=1, run_id=0, stream_base=0):
        assert rec["status"] == 3 and rec["loglike"] == pytest.approx(ref["loglike"], rel=1e-12)


@pytest.mark.parametrize("algorithm", [BPF, APF, RMPF])
def test_observation_times_with_gaps(orc, host_general, algorithm):
    # R/particle_filter_core.R:70-71,124-136: the gap to the previous observation time is that many transitions
    y, ot = sim_y(AR, 6, np.random.default_rng(6)), [1, 2, 4, 7, 8, 12]
    ref = orc.particle_filter(AR, algorithm, 2, 0, 2048, y, THETA[AR], obs_times=ot, seed=3)
    rec, = host_general(AR, algorithm, 2048, y, [THETA[AR]], seed=3, run_id=0, stream_base=0, obs_times=ot)
    check(rec, ref)


@pytest.mark.parametrize("exact", [1, 0])
def test_large_input_path_of_the_cdf_pipeline_scans_the_tile_totals_once(orc, host_general, exact):
    # beyond RS_PREFIX_TILES tiles k_tile_scan reads the scanned totals of k_tile_prefix instead of summing the preceding
    # totals itself: forced here from the second tile on; exact mode reproduces the sequential cumsum whatever the prefix path
    y = sim_y(AR, 5, np.random.default_rng(8))
    thetas = [THETA[AR], list(np.array(THETA[AR]) * 1.05)]
    recs = host_general(AR, BPF, 7000, y, thetas, resample_algorithm=1, exact=exact, prefix_tiles=1)
    for c, rec in enumerate(recs):
        check(rec, orc.particle_filter(AR, 0, 1, 0, 7000, y, thetas[c], seed=77, run_id=1, stream=5 + c), tol=1e-9 if exact else 1e-6)


@pytest.mark.parametrize("algorithm,ralg", [(BPF, 0), (BPF, 2), (RMPF, 1)])
def test_carried_weights_mode(orc, host_general, algorithm, ralg):
    # the deviation mode: weights carried over steps that do not resample (SIS: all of them; SISAR: some; RMPF: none)
    y = sim_y(LG, 7, np.random.default_rng(31))
    rec, = host_general(LG, algorithm, 3000, y, [THETA[LG]], resample_algorithm=ralg, carry=True)
    ref = orc.particle_filter(LG, algorithm, ralg, 0, 3000, y, THETA[LG], seed=77, run_id=1, stream=5, carry_weights=True)
    check(rec, ref)
    if ralg == 0:      # and it is a different estimator from the reference's rule
        plain = orc.particle_filter(LG, algorithm, ralg, 0, 3000, y, THETA[LG], seed=77, run_id=1, stream=5)
        assert abs(plain["loglike"] - ref["loglike"]) > 1e-3
