"""Particle-sharded single filter on the GPU (bssm_filter_run_sharded).  world = 1 runs in the one-GPU suite and
exercises the sharded code path (records, merge kernel, layout descriptors) against the oracle; the two-rank
test needs two GPUs (gpurun --gpus 2) and compares the sharded result with the one-GPU result and the oracle, once
with the exchange fused into the filter kernel through peer memory and once with ncclAllGather (bit-identical)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from bayesssm_b200 import _native as nat
from test_filter_gpu import THETA, sim_y

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AR = 0


def test_world_of_one_matches_oracle(orc, engine):
    from bayesssm_b200 import models, sharding as S
    grp = S.ShardGroup(engine, rank=0, world=1)
    m = models.nonlinear_ar()
    rng = np.random.default_rng(21)
    for N, T, rfn in ((5000, 12, "stratified"), (70001, 8, "systematic")):
        y = sim_y(AR, T, rng)
        ref = orc.particle_filter(AR, 0, 2, 0 if rfn == "stratified" else 1, N, y, THETA[AR], seed=77, stream=5)
        got = S.sharded_bootstrap_filter(y, N, m.init_fn, m.transition_fn, m.log_likelihood_fn, grp, resample_fn=rfn,
                                         precision="f64", seed=77, stream=5, phi=0.8, sigma_x=1.0, sigma_y=0.5)
        assert got["n_resampled"] == ref["n_resampled"] and got["n_local_final"] == N
        assert abs(got["loglike"] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
        np.testing.assert_allclose(got["ess"], ref["ess"], rtol=1e-6)
        np.testing.assert_allclose(got["state_est"], ref["state_est"][:, 0], rtol=1e-6, atol=1e-6)
    grp.close()


WORKER = r'''
import os, sys, json
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from bayesssm_b200 import _native as nat, models, sharding as S
import engine_helpers as eh, oracle
from test_filter_gpu import THETA, sim_y
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = nat.Context(lr)
results = {}
for exch in ("peer", "nccl"):   # the fused peer-memory exchange, then ncclAllGather + k_st_merge: same records, same order, same bits
    grp = S.ShardGroup(ctx, device=torch.device("cuda", lr), exchange=exch)
    assert grp.exchange == exch, (grp.exchange, grp.exchange_note)
    assert bool(ctx.lib.bssm_shard_peer_active(ctx.handle)) == (exch == "peer")
    m = models.nonlinear_ar()
    rng = np.random.default_rng(5)
    out = {}
    for name, N, T, rfn, prec in (("small_f64", 6000, 15, "stratified", "f64"), ("sys_f64", 100003, 10, "systematic", "f64"),
                                  ("big_f32", 1 << 22, 12, "stratified", "f32")):
        y = sim_y(0, T, rng)
        got = S.sharded_bootstrap_filter(y, N, m.init_fn, m.transition_fn, m.log_likelihood_fn, grp, resample_fn=rfn,
                                         precision=prec, seed=31, stream=2, phi=0.8, sigma_x=1.0, sigma_y=0.5)
        counts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(counts, torch.tensor([got["n_local_final"]], dtype=torch.int64, device="cuda"))
        assert sum(int(c) for c in counts) == N, (name, [int(c) for c in counts])
        # every rank holds the same outputs
        ll = torch.tensor([got["loglike"]], dtype=torch.float64, device="cuda")
        lls = [torch.zeros_like(ll) for _ in range(world)]
        dist.all_gather(lls, ll)
        assert all(float(l) == got["loglike"] for l in lls), name
        one = eh.filter_run(ctx, 0, 0, 2, 0 if rfn == "stratified" else 1, N, y, THETA[0], seed=31, stream_base=2,
                            precision=nat.F64 if prec == "f64" else nat.F32, engine=nat.ENGINE_STREAM)
        if prec == "f64":
            ref = oracle.particle_filter(0, 0, 2, 0 if rfn == "stratified" else 1, N, y, THETA[0], seed=31, stream=2)
            assert got["n_resampled"] == ref["n_resampled"] == int(one["n_resampled"][0]), name
            assert abs(got["loglike"] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"]), name
            assert abs(got["loglike"] - one["loglike"][0]) <= 1e-9 * abs(ref["loglike"]), name
            np.testing.assert_allclose(got["ess"], ref["ess"], rtol=1e-6)
            np.testing.assert_allclose(got["state_est"], ref["state_est"][:, 0], rtol=1e-6, atol=1e-6)
        else:
            assert abs(got["loglike"] - one["loglike"][0]) < 0.05, (name, got["loglike"], one["loglike"][0])
            np.testing.assert_allclose(got["state_est"], one["state_est"][0][:, 0], atol=0.02)
        out[name] = got["loglike"]
    # Kalman gate on the sharded path: linear-Gaussian model, SISR, throughput precision, a few seeds
    lg = models.linear_gaussian()
    rng2 = np.random.default_rng(21)
    ylg = sim_y(1, 60, rng2)
    exact = oracle.kalman_loglik(ylg, 0.8, 1.0, 1.0)
    lls = np.array([S.sharded_bootstrap_filter(ylg, 1 << 21, lg.init_fn, lg.transition_fn, lg.log_likelihood_fn, grp,
                                               resample_algorithm="SISR", precision="f32", seed=200 + s, phi=0.8, sigma_x=1.0,
                                               sigma_y=1.0)["loglike"] for s in range(5)])
    se = lls.std(ddof=1) / np.sqrt(len(lls))
    assert abs(lls.mean() - exact) < 3 * se + 5e-3, (lls, exact)
    out["kalman_diff"] = float(lls.mean() - exact)
    # capacity overflow is reported, not a crash: a very sharp likelihood piles the offspring on one rank
    y = np.array([0.3, 0.1]); 
    try:
        S.sharded_bootstrap_filter(y, 1 << 16, m.init_fn, m.transition_fn, m.log_likelihood_fn, grp, resample_algorithm="SISR",
                                   precision="f64", seed=3, capacity_factor=1.0, phi=0.8, sigma_x=1.0, sigma_y=1e-3)
        overflow = False
    except nat.EngineError as e:
        overflow = e.status == nat.ERR_CAPACITY
    flags = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(flags, torch.tensor([int(overflow)], dtype=torch.int64, device="cuda"))
    assert len({int(f) for f in flags}) == 1, "ranks disagree on the capacity overflow"
    results[exch] = dict(out, overflow=overflow)
    grp.close()
assert results["peer"] == results["nccl"], results
# a peer that never arrives is an error after the timeout, not a hung GPU: rank 1 stays out of one collective call
os.environ["BSSM_PEER_TIMEOUT_MS"] = "400"
grp = S.ShardGroup(ctx, device=torch.device("cuda", lr), exchange="peer")
timed_out = True
if rank == 0:
    try:
        S.sharded_bootstrap_filter(sim_y(0, 6, np.random.default_rng(1)), 6000, m.init_fn, m.transition_fn, m.log_likelihood_fn, grp,
                                   precision="f64", seed=1, phi=0.8, sigma_x=1.0, sigma_y=0.5)
        timed_out = False
    except nat.EngineError as e:
        timed_out = e.status == nat.ERR_NCCL
assert timed_out
dist.barrier()
grp.close()
del os.environ["BSSM_PEER_TIMEOUT_MS"]
# ... and a fresh group works again
grp = S.ShardGroup(ctx, device=torch.device("cuda", lr), exchange="peer")
again = S.sharded_bootstrap_filter(sim_y(0, 15, np.random.default_rng(5)), 6000, m.init_fn, m.transition_fn, m.log_likelihood_fn, grp,
                                   precision="f64", seed=31, stream=2, phi=0.8, sigma_x=1.0, sigma_y=0.5)
assert again["loglike"] == results["peer"]["small_f64"], (again["loglike"], results["peer"]["small_f64"])
grp.close()
ctx.close()
dist.barrier(); dist.destroy_process_group()
print("ok", rank, json.dumps(out), "overflow", overflow)
'''


def test_two_ranks_match_one_gpu_and_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
