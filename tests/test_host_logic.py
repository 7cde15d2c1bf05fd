"""Host-side logic that needs no GPU: argument validation mirroring the reference's error behaviour,
diagnostics (R/ESS.R, R/rhat.R), chain sharding and the gather (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import bayesssm_b200 as b
from bayesssm_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_tune_control_defaults_and_validation():
    # tests/testthat/test-pmmh.R:5-74
    tc = b.default_tune_control()
    assert tc["pilot_proposal_sd"] == 0.5 and tc["pilot_n"] == 100 and tc["pilot_m"] == 2000
    assert tc["pilot_target_var"] == 1 and tc["pilot_burn_in"] == 500 and tc["pilot_reps"] == 100
    assert tc["pilot_resample_algorithm"] == "SISAR" and tc["pilot_resample_fn"] == "stratified"
    for bad in (dict(pilot_n=0), dict(pilot_m=-1), dict(pilot_proposal_sd=-0.1), dict(pilot_reps=0),
                dict(pilot_resample_algorithm="bogus"), dict(pilot_resample_fn="bogus")):
        with pytest.raises(ValueError):
            b.default_tune_control(**bad)


def test_filter_argument_validation_before_any_device_work():
    # tests/testthat/test-particle_filter_core.R:19-107 (the checks that do not depend on closures)
    m = b.models.nonlinear_ar()
    ok = dict(phi=0.8, sigma_x=1.0, sigma_y=0.5)
    args = (m.init_fn, m.transition_fn, m.log_likelihood_fn)
    with pytest.raises(ValueError, match="num_particles"):
        b.bootstrap_filter(np.zeros(5), 0, *args, **ok)
    with pytest.raises(ValueError, match="'y'"):
        b.bootstrap_filter(np.array([0.0, np.nan]), 10, *args, **ok)
    with pytest.raises(ValueError, match="obs_times"):
        b.bootstrap_filter(np.zeros(5), 10, *args, obs_times=[1, 2, 3], **ok)
    with pytest.raises(ValueError, match="obs_times"):
        b.bootstrap_filter(np.zeros(3), 10, *args, obs_times=[1, 2.5, 3], **ok)
    with pytest.raises(ValueError, match="obs_times"):
        b.bootstrap_filter(np.zeros(3), 10, *args, obs_times=[3, 2, 1], **ok)
    with pytest.raises(ValueError, match="resample_fn"):
        b.bootstrap_filter(np.zeros(3), 10, *args, resample_fn="bogus", **ok)
    with pytest.raises(TypeError, match="device-model"):
        b.bootstrap_filter(np.zeros(3), 10, lambda n: np.zeros(n), m.transition_fn, m.log_likelihood_fn, **ok)
    with pytest.raises(ValueError, match="different device models"):
        b.bootstrap_filter(np.zeros(3), 10, b.models.linear_gaussian().init_fn, m.transition_fn, m.log_likelihood_fn, **ok)


def test_pmmh_argument_validation():
    # tests/testthat/test-pmmh.R:84-361 (name matching, burn_in, transforms, chain count)
    m = b.models.nonlinear_ar()
    pri = {"phi": b.priors.uniform(0, 1), "sigma_x": b.priors.exponential(1), "sigma_y": b.priors.exponential(1)}
    init = [{"phi": 0.8, "sigma_x": 1.0, "sigma_y": 0.5}] * 2
    base = dict(pf_wrapper=b.bootstrap_filter, y=np.zeros(5), m=10, init_fn=m.init_fn, transition_fn=m.transition_fn,
                log_likelihood_fn=m.log_likelihood_fn, log_priors=pri, pilot_init_params=init, burn_in=2, num_chains=2)
    with pytest.raises(ValueError, match="burn_in"):
        b.pmmh(**{**base, "burn_in": 10})
    with pytest.raises(ValueError, match="pilot_init_params"):
        b.pmmh(**{**base, "num_chains": 3})
    with pytest.raises(ValueError, match="do not match"):
        b.pmmh(**{**base, "log_priors": {"phi": pri["phi"]}})
    with pytest.raises(ValueError, match="param_transform must include"):
        b.pmmh(**{**base, "param_transform": {"phi": "logit"}})
    with pytest.raises(ValueError, match="param_transform must be a list"):
        b.pmmh(**{**base, "param_transform": "log"})
    with pytest.raises(TypeError, match="pf_wrapper"):
        b.pmmh(**{**base, "pf_wrapper": print})


def test_shard_chains_partition():
    for n, w in ((1024, 8), (10, 4), (7, 2), (3, 3)):
        parts = [D.shard_chains(n, r, w) for r in range(w)]
        assert sum(c for _, c in parts) == n
        assert parts[0][0] == 0 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_chain_gather_world_size_2_gloo(tmp_path):
    """Two CPU processes (gloo): each 'runs' its shard with a stub keyed by the GLOBAL chain id, the gathered
    arrays come back in global chain order on both ranks."""
    script = tmp_path / "worker.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch.distributed as dist
from bayesssm_b200 import distributed as D
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
num_chains, m, p = 5, 4, 3
def run_local(base, count):
    ids = np.arange(base, base + count)
    return {{"theta_chain": (ids[:, None, None] * 100 + np.arange(m)[None, :, None] * 10 + np.arange(p)[None, None, :]).astype(np.float64),
            "loglike_chain": -ids[:, None] * np.ones((1, m)), "n_accept": ids.astype(np.int32) * 2,
            "target_n": np.full(count, 50, dtype=np.int32), "status": np.zeros(count, dtype=np.int32)}}
out = D.pmmh_sharded(run_local, num_chains, rank, world)
full = run_local(0, num_chains)
for k in full:
    assert out[k].shape == full[k].shape and np.array_equal(out[k], full[k]), k
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_pmmh_output_summary_and_print():
    """tests/testthat/test-summary.R:1-27 (summary.pmmh_output, R/summary.R:28-54) and print.pmmh_output
    (R/print.R:30-66): dummy chains, `chain` as the LAST column as in the reference's test."""
    import pandas as pd
    from bayesssm_b200.pmmh import PmmhOutput
    rng = np.random.default_rng(1)
    chains = pd.concat([pd.DataFrame({"param1": rng.standard_normal(100), "param2": rng.standard_normal(100), "chain": c})
                        for c in (1, 2)], ignore_index=True)
    out = PmmhOutput(theta_chain=chains, diagnostics={"ess": {"param1": 200.9, "param2": 190.0},
                                                      "rhat": {"param1": 1.0104, "param2": 1.0}})
    sm = out.summary()
    assert list(sm.columns) == ["mean", "sd", "median", "2.5%", "97.5%", "ESS", "Rhat"]
    assert list(sm.index) == ["param1", "param2"]
    assert sm.loc["param1", "ESS"] == 200.9 and sm.loc["param1", "Rhat"] == 1.0104
    x = chains["param2"].to_numpy()
    np.testing.assert_allclose(sm.loc["param2", ["mean", "sd", "median"]].to_numpy(dtype=float),
                               [x.mean(), x.std(ddof=1), np.median(x)])
    xs = np.sort(x)                                        # quantile type 7: h = (n - 1) q, linear between order statistics
    h = (len(xs) - 1) * 0.025
    np.testing.assert_allclose(sm.loc["param2", "2.5%"], xs[int(h)] + (h - int(h)) * (xs[int(h) + 1] - xs[int(h)]))
    text = str(out).splitlines()
    assert text[0] == "PMMH Results Summary:"
    assert text[1].split() == ["Parameter", "Mean", "SD", "Median", "2.5%", "97.5%", "ESS", "Rhat"]
    row = text[2].split()
    assert row[0] == "param1" and row[6] == "200" and float(row[7]) == 1.01      # ESS floored, Rhat to 3 digits
    assert float(row[1]) == round(chains["param1"].mean(), 2)
