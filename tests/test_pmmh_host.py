"""The device-resident PMMH without a GPU: the kernel text of bayesssm_b200/csrc/bssm_pmmh.cuh (proposal, accept /
reject, pilot statistics, tuning ...; one thread per chain) driven in bssm_pmmh_run()'s order, with every batched filter
pass run by the persistent kernel's text, all over the SIMT emulation (tests/host_pmmh.cpp, tests/simt_emu.h).  Compared
chain by chain, draw by draw, with the oracle restating R/pmmh.R:345-505 and R/pmmh_tuning.R:29-64,111-317 on the same
Philox streams -- the comparison tests/test_pmmh_gpu.py makes on the device (rtol 1e-8 there; 1e-10 here)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRIOR = dict(prior_kind=[3, 2, 2], prior_a=[0.0, 1.0, 1.0], prior_b=[1.0, 0.0, 0.0])   # README.md:153-168
KEYS = ("pilot_theta_chain", "pilot_loglike_chain", "pilot_theta_mean", "pilot_theta_cov", "pilot_loglikes", "proposal_chol",
        "theta_chain", "loglike_chain")


@pytest.fixture(scope="module")
def host_pmmh(tmp_path_factory):
    exe = tmp_path_factory.mktemp("hp") / "host_pmmh"
    subprocess.run(["g++", "-O1", "-std=c++20", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_pmmh.cpp")], check=True)

    def run(y, inits, transform, sd, pilot_n, pilot_m, pilot_reps, m, seed, chain_id_base, pilot_ralg=2, pilot_rfn=0, fixed_n=0, G=2):
        inits = np.ascontiguousarray(inits, dtype=np.float64)
        C = len(inits)
        cfg = np.array(PRIOR["prior_kind"] + PRIOR["prior_a"] + PRIOR["prior_b"] + list(transform) + list(sd), dtype=np.float64)
        args = [C, len(y), pilot_n, pilot_m, pilot_reps, m, seed, chain_id_base, pilot_ralg, pilot_rfn, fixed_n, G]
        r = subprocess.run([str(exe)] + [str(a) for a in args], input=np.asarray(y, dtype=np.float64).tobytes() + inits.tobytes() + cfg.tobytes(),
                           capture_output=True, timeout=900)
        assert r.returncode == 0, r.stderr.decode()[-2000:]
        lines, out = r.stdout.decode().strip().splitlines(), []
        for c in range(C):
            blk = lines[c * 9:(c + 1) * 9]
            h = blk[0].split()
            rec = {"status": int(h[3]), "target_n": int(h[5]), "n_accept": int(h[7])}
            rec.update({ln.split()[0]: np.array(ln.split()[1:], float) for ln in blk[1:]})
            out.append(rec)
        return out
    return run


def readme_data(T, rng):
    x, ys = rng.standard_normal(), []
    for _ in range(T):
        x = 0.8 * x + np.sin(x) + rng.standard_normal()
        ys.append(x + 0.5 * rng.standard_normal())
    return np.array(ys)


@pytest.mark.parametrize("transform,pilot_ralg,pilot_rfn,G", [([2, 1, 1], 2, 0, 2), ([0, 0, 0], 1, 1, 3)])
def test_emulated_pmmh_reproduces_the_oracle_chains(orc, host_pmmh, transform, pilot_ralg, pilot_rfn, G):
    y = readme_data(12, np.random.default_rng(1405))
    inits = np.array([[0.8, 1.0, 0.5], [0.5, 0.7, 1.2], [0.3, 1.5, 0.8]])
    kw = dict(transform=transform, pilot_proposal_sd=[0.1, 0.15, 0.2], pilot_n=64, pilot_m=30, pilot_reps=6, m=40, seed=99)
    got = host_pmmh(y, inits, transform, kw["pilot_proposal_sd"], 64, 30, 6, 40, 99, 4, pilot_ralg, pilot_rfn, G=G)
    targets = set()
    for c, rec in enumerate(got):
        ref = orc.pmmh_chain(0, 0, y, inits[c], chain_id=4 + c, pilot_resample_algorithm=pilot_ralg, pilot_resample_fn=pilot_rfn,
                             **PRIOR, **kw)
        assert ref["status"] == 0 and rec["status"] == 0
        assert rec["target_n"] == ref["target_n"] and rec["n_accept"] == ref["n_accept"]
        for k in KEYS:
            np.testing.assert_allclose(rec[k], np.asarray(ref[k]).ravel(), rtol=1e-10, atol=1e-12, err_msg=k)
        targets.add(rec["target_n"])
    assert len(targets) > 1          # the tuned particle counts differ between the chains: the main phase ran a ragged batch


def test_invalid_start_stops_that_chain_only(host_pmmh):
    # R/pmmh_tuning.R:135-143: "Initial parameter values are invalid" -> per-chain status BSSM_ERR_PRIOR_INIT (5)
    y = readme_data(8, np.random.default_rng(1))
    got = host_pmmh(y, [[0.8, 1.0, 0.5], [1.5, 1.0, 0.5]], [0, 0, 0], [0.1] * 3, 32, 10, 4, 10, 1, 0)
    assert got[0]["status"] == 0 and got[1]["status"] == 5
    assert np.isfinite(got[0]["theta_chain"]).all()
