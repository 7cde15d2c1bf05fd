"""The streaming bootstrap-filter engine without a GPU: the kernel text of bayesssm_b200/csrc/bssm_stream.cuh is
compiled by g++ over a small SIMT emulation (tests/simt_emu.h: the threads of a block are fibers, barriers and warp
shuffles are real rendezvous, shared memory is per block, blocks run in a chosen order) and driven like
stream_launch() drives it (tests/host_stream.cpp) -- also in its particle-sharded form, with 2 - 4 emulated ranks
and the per-observation all-gather of the records as a memcpy.  Results are compared with the oracle's Philox-mode
filter (R/particle_filter_core.R:76-266 + src/resampling.cpp:16-66 restated in oracle/pf_oracle.c).

In the parity precision (f64) everything but the summation order and the device's FMA contraction is reproduced, so
the tolerance is 1e-9 where the GPU tests allow 1e-6.  The throughput precision (f32) runs with libm in place of the
SFU approximations and is held to the statistical tolerance of its GPU tests."""
import os
import subprocess

import numpy as np
import pytest

from test_filter_gpu import THETA, sim_y

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AR, LG, RWD, ARCOS = 0, 1, 2, 4


@pytest.fixture(scope="module")
def host_stream(tmp_path_factory):
    exe = tmp_path_factory.mktemp("hs") / "host_stream"
    subprocess.run(["g++", "-O1", "-std=c++20", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_stream.cpp")], check=True)

    def run(model, N, y, thetas, precision=64, threads=256, bpc=2, resample_fn=0, resample_algorithm=2, threshold=-1.0,
            seed=1405, run_id=2, stream_base=3, world=1, capacity_factor=1.5, block_order=0, n_per=(), obs_times=None, chain=0, mn_ahead=False):
        y = np.ascontiguousarray(y, dtype=np.float64)
        th = np.zeros((len(thetas), 3))
        for c, t in enumerate(thetas):
            th[c, :len(t)] = t
        args = [model, precision, threads, N, len(y), len(thetas), bpc, resample_fn, resample_algorithm, threshold, seed, run_id,
                stream_base, world, capacity_factor, block_order] + list(n_per)
        env = dict(os.environ)
        if obs_times is not None:
            env["EMU_OBS_TIMES"] = ",".join(str(int(t)) for t in obs_times)
        if mn_ahead:   # multinomial: position arrays doubled by the observation's parity and laid out for every observation
            env["EMU_MN_AHEAD"] = "1"
        if chain:      # the chain-persistent kernel k_st_chain: cooperative launches over groups of `chain` filters
            env["EMU_CHAIN"] = str(chain)
        r = subprocess.run([str(exe)] + [str(a) for a in args], input=y.tobytes() + th.tobytes(), capture_output=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        lines, recs = r.stdout.decode().strip().splitlines(), []
        for i in range(0, len(lines), 4):
            h = lines[i].split()
            recs.append({"rank": int(h[1]), "filter": int(h[3]), "loglike": float(h[5]), "n_resampled": int(h[7]),
                         "status": int(h[9]), "early_exit": int(h[11]),
                         "ess": np.array(lines[i + 1].split()[1:], float), "state_est": np.array(lines[i + 2].split()[1:], float),
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


# sizes around the tile boundaries (f64 tile = 1024 / 512 particles), both block sizes, both resamplers, all block orders
@pytest.mark.parametrize("N,T,threads,bpc,rfn,order", [(1, 5, 128, 1, 0, 0), (3, 6, 256, 1, 1, 0), (1024, 6, 256, 1, 0, 0),
                                                       (1025, 6, 128, 3, 1, 1), (3000, 8, 256, 2, 0, 2), (70001, 4, 256, 5, 0, 2)])
def test_f64_kernel_text_matches_oracle(orc, host_stream, N, T, threads, bpc, rfn, order):
    y = sim_y(AR, T, np.random.default_rng(N))
    ref = orc.particle_filter(AR, 0, 2, rfn, N, y, THETA[AR], seed=1405, run_id=2, stream=3)
    rec, = host_stream(AR, N, y, [THETA[AR]], threads=threads, bpc=bpc, resample_fn=rfn, block_order=order)
    check(rec, ref)


@pytest.mark.parametrize("model,ralg", [(LG, 2), (RWD, 0), (ARCOS, 1)])
def test_models_resample_algorithms_and_batches(orc, host_stream, model, ralg):
    y = sim_y(AR if model == ARCOS else model, 6, np.random.default_rng(5))
    base = np.array(THETA[AR] if model == ARCOS else THETA[model])
    thetas = [list(base * (1 + 0.05 * c)) for c in range(3)]
    recs = host_stream(model, 2000, y, thetas, threads=128, bpc=4, resample_algorithm=ralg, seed=9, run_id=0, stream_base=1)
    assert [r["filter"] for r in recs] == [0, 1, 2]
    for c, rec in enumerate(recs):     # filter c: its own theta, Philox stream stream_base + c
        check(rec, orc.particle_filter(model, 0, ralg, 0, 2000, y, thetas[c], seed=9, stream=1 + c))


def test_ragged_batch_particle_counts_per_filter(orc, host_stream):
    # PMMH's tuned target_n differs from chain to chain (R/pmmh_tuning.R:54-57): FilterDev::n_per
    y = sim_y(AR, 6, np.random.default_rng(3))
    ns = [3000, 50, 1777]
    thetas = [list(np.array(THETA[AR]) * (1 + 0.05 * c)) for c in range(3)]
    recs = host_stream(AR, 3000, y, thetas, threads=128, bpc=3, seed=5, run_id=1, stream_base=2, block_order=2, n_per=ns)
    for c, rec in enumerate(recs):
        check(rec, orc.particle_filter(AR, 0, 2, 0, ns[c], y, thetas[c], seed=5, run_id=1, stream=2 + c))


def test_threshold_early_exit_and_no_observations(orc, host_stream):
    y = sim_y(AR, 6, np.random.default_rng(6))
    rec, = host_stream(AR, 2048, y, [THETA[AR]], threshold=1500.0, seed=3, run_id=0, stream_base=0)
    check(rec, orc.particle_filter(AR, 0, 2, 0, 2048, y, THETA[AR], seed=3, threshold=1500.0))
    y2, th = np.array([0.1, 1e6, 0.2]), [0.8, 1.0, 1e-3]               # R/particle_filter_core.R:189-202
    rec, = host_stream(AR, 512, y2, [th], bpc=1, seed=3, run_id=0, stream_base=0)
    ref = orc.particle_filter(AR, 0, 2, 0, 512, y2, th, seed=3)
    assert ref["early_exit"] == 1 and rec["early_exit"] == 1 and rec["loglike"] == -np.inf and rec["status"] == 0
    np.testing.assert_allclose(rec["ess"], ref["ess"], rtol=1e-9)
    rec, = host_stream(AR, 777, np.zeros(0), [THETA[AR]], bpc=1, seed=3, run_id=0, stream_base=0)
    ref = orc.particle_filter(AR, 0, 2, 0, 777, np.zeros(0), THETA[AR], seed=3)
    assert rec["ess"][0] == 777 and abs(rec["state_est"][0] - ref["state_est"][0, 0]) < 1e-12


def test_degenerate_weights_a_few_particles_take_everything(orc, host_stream):
    # a very sharp likelihood: heavy sources spanning many output chunks, empty tiles elsewhere
    y, th = sim_y(AR, 5, np.random.default_rng(11)), [0.8, 1.0, 2e-4]
    rec, = host_stream(AR, 50000, y, [th], bpc=4, resample_algorithm=1, seed=13, run_id=0, stream_base=0, block_order=2)
    check(rec, orc.particle_filter(AR, 0, 1, 0, 50000, y, th, seed=13))


@pytest.mark.parametrize("N,world,threads,rfn,ralg", [(4100, 2, 256, 0, 2), (3001, 3, 128, 1, 1), (40000, 4, 128, 0, 2)])
def test_particle_sharded_ranks_reproduce_the_one_gpu_filter(orc, host_stream, N, world, threads, rfn, ralg):
    """bssm_shard.cu's per-observation sequence (k_st_step, all-gather of the 64-byte records, k_st_merge,
    k_st_resample on each rank's own offspring) with emulated ranks: every rank reports the same numbers, and they
    are the unsharded filter's."""
    y = sim_y(AR, 5, np.random.default_rng(N))
    ref = orc.particle_filter(AR, 0, ralg, rfn, N, y, THETA[AR], seed=1405, run_id=2, stream=3)
    recs = host_stream(AR, N, y, [THETA[AR]], threads=threads, resample_fn=rfn, resample_algorithm=ralg, world=world, block_order=2)
    assert [r["rank"] for r in recs] == list(range(world))
    for rec in recs:
        check(rec, ref)
        assert rec["loglike"] == recs[0]["loglike"] and np.array_equal(rec["state_est"], recs[0]["state_est"])


def test_capacity_overflow_is_reported_by_every_rank(host_stream):
    y, th = sim_y(AR, 5, np.random.default_rng(11)), [0.8, 1.0, 2e-4]   # nearly all offspring belong to one rank
    recs = host_stream(AR, 20000, y, [th], threads=128, bpc=4, resample_algorithm=1, seed=13, world=4, capacity_factor=1.0)
    assert [r["status"] for r in recs] == [10, 10, 10, 10]               # BSSM_ERR_CAPACITY


def test_capacity_factor_equal_to_the_rank_count_never_overflows(orc, host_stream):
    # storage for N particles per rank: even when one rank inherits (nearly) every offspring the run completes
    y, th = sim_y(AR, 5, np.random.default_rng(11)), [0.8, 1.0, 2e-4]
    ref = orc.particle_filter(AR, 0, 1, 0, 20000, y, th, seed=13)
    for rec in host_stream(AR, 20000, y, [th], threads=128, bpc=4, resample_algorithm=1, seed=13, run_id=0, stream_base=0, world=4,
                           capacity_factor=4.0, block_order=2):
        check(rec, ref)


def test_kernels_stay_inside_their_buffers_under_address_sanitizer(tmp_path):
    """Both harnesses allocate exactly what stream_launch() / fast_launch() allocate (particle rows, records, prefix arrays,
    LL buffers, dynamic shared memory); AddressSanitizer then sees any access past them."""
    y = sim_y(AR, 5, np.random.default_rng(11))
    runs = {"host_stream": [([0.8, 1.0, 2e-4], [0, 64, 256, 4099, 5, 1, 4, 0, 1, -1.0, 13, 0, 0, 2, 2.0, 2]),
                            ([0.8, 1.0, 0.5], [0, 64, 128, 5001, 5, 1, 3, 1, 2, -1.0, 13, 0, 0, 3, 3.0, 1]),
                            ([0.8, 1.0, 0.5], [0, 32, 256, 3001, 5, 1, 2, 0, 2, -1.0, 13, 0, 0, 1, 1.5, 0])],
            "host_fast": [([0.8, 1.0, 2e-4], [0, 0, 4, 1, 3000, 5, 1, 0, 1, -1.0, 13, 0, 0]),
                          ([0.8, 1.0, 2e-4], [0, 1, 3, 1, 7000, 5, 1, 0, 1, -1.0, 13, 0, 0]),
                          ([0.8, 1.0, 0.5], [0, 3, 2, 1, 5001, 5, 1, 0, 2, -1.0, 13, 0, 0]),
                          ([0.8, 1.0, 0.5], [0, 2, 5, 1, 2049, 5, 1, 1, 1, -1.0, 13, 0, 0])]}
    builds = {name: subprocess.Popen(["g++", "-O1", "-fsanitize=address", "-std=c++20", "-ffp-contract=off", "-Wno-unknown-pragmas",
                                      "-pthread", "-o", str(tmp_path / (name + "_asan")), os.path.join(ROOT, "tests", name + ".cpp")])
              for name in runs}                                    # both compilations at once
    for name, cases in runs.items():
        assert builds[name].wait() == 0
        for th, args in cases:
            r = subprocess.run([str(tmp_path / (name + "_asan"))] + [str(a) for a in args], input=y.tobytes() + np.array([th]).tobytes(),
                               capture_output=True, timeout=900)
            assert r.returncode == 0 and b"ERROR: AddressSanitizer" not in r.stderr, r.stderr.decode()[-3000:]
            assert b"status 0" in r.stdout


def test_f32_kernel_text_is_close_and_rank_independent(orc, host_stream):
    y = sim_y(AR, 5, np.random.default_rng(6000))
    ref = orc.particle_filter(AR, 0, 2, 0, 6000, y, THETA[AR], seed=1405, run_id=2, stream=3)
    rec, = host_stream(AR, 6000, y, [THETA[AR]], precision=32)
    assert rec["status"] == 0 and rec["n_resampled"] == ref["n_resampled"]
    assert abs(rec["loglike"] - ref["loglike"]) < 5e-3 and np.abs(rec["state_est"] - ref["state_est"][:, 0]).max() < 5e-3
    two = host_stream(AR, 6000, y, [THETA[AR]], precision=32, world=2, threads=128, block_order=1)
    assert two[0]["loglike"] == two[1]["loglike"] and abs(two[0]["loglike"] - ref["loglike"]) < 5e-2


def test_nan_observation_is_reported_as_in_the_oracle(orc, host_stream):
    # NaN weights: R's `if (all(lw < -1e8))` raises "missing value where TRUE/FALSE needed" (R/particle_filter_core.R:189);
    # here status 3 (BSSM_ERR_NAN_WEIGHT), the log-likelihood staying at the last finite observation
    y = sim_y(AR, 5, np.random.default_rng(1))
    y[2] = np.nan
    ref = orc.particle_filter(AR, 0, 2, 0, 3000, y, THETA[AR], seed=1)
    assert ref["status"] == 3
    for rec in host_stream(AR, 3000, y, [THETA[AR]], threads=128, seed=1, run_id=0, stream_base=0, world=3, capacity_factor=3.0, block_order=2):
        assert rec["status"] == 3 and rec["loglike"] == pytest.approx(ref["loglike"], rel=1e-12)


def test_observation_times_with_gaps(orc, host_stream):
    # R/particle_filter_core.R:70-71,124-136: the gap to the previous observation time is that many transitions
    y, ot = sim_y(AR, 6, np.random.default_rng(6)), [1, 2, 4, 7, 8, 12]
    ref = orc.particle_filter(AR, 0, 2, 0, 2048, y, THETA[AR], obs_times=ot, seed=3)
    for rec in host_stream(AR, 2048, y, [THETA[AR]], seed=3, run_id=0, stream_base=0, world=2, capacity_factor=2.0, block_order=1, obs_times=ot):
        check(rec, ref)


# multinomial resampling on the streaming engine: sorted uniforms from exponential spacings (k_st_mn_sums / _scan / _positions),
# offspring ranges counted in the staged positions.  Oracle: orc_resample_multinomial_sorted inside the Philox-mode filter
# (resample_fn = 3).  Sizes around the tile edges, several blocks per filter, a batch, degenerate weights (several chunks)
@pytest.mark.parametrize("N,T,threads,bpc,ralg", [(1, 4, 128, 1, 1), (37, 6, 128, 1, 1), (1024, 6, 256, 1, 2), (1025, 6, 128, 2, 1),
                                                  (5000, 8, 128, 3, 2), (20011, 5, 256, 4, 1)])
def test_multinomial_by_sorted_uniforms_matches_the_oracle(orc, host_stream, N, T, threads, bpc, ralg):
    y = sim_y(AR, T, np.random.default_rng(N))
    ref = orc.particle_filter(AR, 0, ralg, 3, N, y, THETA[AR], seed=1405, run_id=2, stream=3)
    rec, = host_stream(AR, N, y, [THETA[AR]], threads=threads, bpc=bpc, resample_fn=2, resample_algorithm=ralg)
    check(rec, ref)
    # the layout of the one-GPU launcher: positions of every observation, buffers doubled by the observation's parity
    rec, = host_stream(AR, N, y, [THETA[AR]], threads=threads, bpc=bpc, resample_fn=2, resample_algorithm=ralg, mn_ahead=True)
    check(rec, ref)


def test_multinomial_batch_degenerate_weights_and_f32(orc, host_stream):
    y = sim_y(AR, 5, np.random.default_rng(8))
    thetas = [THETA[AR], [0.8, 1.0, 2e-3], [0.6, 1.2, 0.7]]      # the second filter: a few particles take everything
    recs = host_stream(AR, 6000, y, thetas, threads=128, bpc=2, resample_fn=2, resample_algorithm=1, seed=4, run_id=0, stream_base=0)
    for c, rec in enumerate(recs):
        check(rec, orc.particle_filter(AR, 0, 1, 3, 6000, y, thetas[c], seed=4, stream=c))
    rec, = host_stream(AR, 20000, y, [THETA[AR]], precision=32, threads=128, bpc=3, resample_fn=2, resample_algorithm=2)
    ref = orc.particle_filter(AR, 0, 2, 3, 20000, y, THETA[AR], seed=1405, run_id=2, stream=3)
    assert rec["status"] == 0 and rec["n_resampled"] == ref["n_resampled"]
    assert abs(rec["loglike"] - ref["loglike"]) < 2e-2 and np.abs(rec["state_est"] - ref["state_est"][:, 0]).max() < 2e-2


# ---- the chain-persistent kernel (k_st_chain): all observations in one cooperative launch per group of filters ----
@pytest.mark.parametrize("N,T,threads,bpc,rfn,ralg,group", [(3000, 8, 128, 3, 0, 2, 2), (1025, 6, 256, 1, 1, 1, 5), (7000, 7, 128, 5, 0, 2, 1),
                                                            (2048, 6, 256, 2, 0, 0, 3)])
def test_chain_persistent_kernel_matches_oracle(orc, host_stream, N, T, threads, bpc, rfn, ralg, group):
    y = sim_y(AR, T, np.random.default_rng(N + 1))
    thetas = [list(np.array(THETA[AR]) * (1 + 0.04 * c)) for c in range(3)]
    recs = host_stream(AR, N, y, thetas, threads=threads, bpc=bpc, resample_fn=rfn, resample_algorithm=ralg, seed=21, run_id=1,
                       stream_base=4, chain=group)
    for c, rec in enumerate(recs):
        check(rec, orc.particle_filter(AR, 0, ralg, rfn, N, y, thetas[c], seed=21, run_id=1, stream=4 + c))


def test_chain_persistent_kernel_ragged_counts_gaps_early_exit_and_f32(orc, host_stream):
    y = sim_y(AR, 6, np.random.default_rng(3))
    ns = [3000, 50, 1777]
    thetas = [list(np.array(THETA[AR]) * (1 + 0.05 * c)) for c in range(3)]
    recs = host_stream(AR, 3000, y, thetas, threads=128, bpc=3, seed=5, run_id=1, stream_base=2, n_per=ns, chain=3)
    for c, rec in enumerate(recs):
        check(rec, orc.particle_filter(AR, 0, 2, 0, ns[c], y, thetas[c], seed=5, run_id=1, stream=2 + c))
    # observation times with gaps
    times = [1, 2, 5, 6, 9, 10]
    recs = host_stream(AR, 2500, y, thetas[:2], threads=128, bpc=2, seed=8, run_id=0, stream_base=0, obs_times=times, chain=2)
    for c, rec in enumerate(recs):
        check(rec, orc.particle_filter(AR, 0, 2, 0, 2500, y, thetas[c], seed=8, stream=c, obs_times=times))
    # one filter of the batch dies (all weights below -1e8), the other carries on
    y2 = np.array([0.1, 1e6, 0.2])
    th_dead, th_ok = [0.8, 1.0, 1e-3], [0.8, 1.0, 1e7]
    recs = host_stream(AR, 1500, y2, [th_dead, th_ok], threads=128, bpc=2, seed=3, run_id=0, stream_base=0, chain=2)
    ref = [orc.particle_filter(AR, 0, 2, 0, 1500, y2, th, seed=3, stream=c) for c, th in enumerate([th_dead, th_ok])]
    assert ref[0]["early_exit"] == 1 and recs[0]["early_exit"] == 1 and recs[0]["loglike"] == -np.inf
    check(recs[1], ref[1])
    # throughput precision: the same kernel text in f32 agrees with the launch-per-body form bit for bit
    a = host_stream(AR, 5000, y, thetas, precision=32, threads=128, bpc=4, seed=5, run_id=1, stream_base=2)
    b = host_stream(AR, 5000, y, thetas, precision=32, threads=128, bpc=4, seed=5, run_id=1, stream_base=2, chain=3)
    for ra, rb in zip(a, b):
        assert ra["loglike"] == rb["loglike"] and ra["n_resampled"] == rb["n_resampled"]
        np.testing.assert_array_equal(ra["state_est"], rb["state_est"])
