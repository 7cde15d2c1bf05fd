"""MCMC diagnostics without a GPU: (1) the numpy restatement of R/ESS.R and R/rhat.R (oracle/mcmc_diag.py) against
the reference's own tests for them (tests/testthat/test-ESS.R, test-rhat.R); (2) the per-thread bodies of the
device kernels (bayesssm_b200/csrc/bssm_diag.cuh), compiled for the host and run thread by thread, against that
restatement; (3) the input checks of the public ess() / rhat(), which raise before anything reaches the device."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mcmc_diag as od  # noqa: E402

import bayesssm_b200 as b  # noqa: E402


def ar1(rng, m, k, rho):
    x = np.zeros((m, k))
    x[0] = rng.standard_normal(k)
    for t in range(1, m):
        x[t] = rho * x[t - 1] + rng.standard_normal(k)
    return x


def test_oracle_against_the_reference_tests():
    rng = np.random.default_rng(1405)
    iid = rng.standard_normal((1000, 3))
    assert abs(od.ess_matrix(iid) - 3000) < 0.05 * 3000                 # test-ESS.R:1-5
    assert od.ess_matrix(ar1(rng, 1000, 3, 0.9)) < 3000                 # test-ESS.R:7-22
    with pytest.raises(ValueError, match="Number of iterations must be at least 2"):
        od.ess_matrix(rng.standard_normal((1, 3)))                      # test-ESS.R:43-46
    with pytest.raises(ValueError, match="Number of chains must be at least 2"):
        od.ess_matrix(rng.standard_normal((6, 1)))                      # test-ESS.R:48-51
    assert np.isnan(od.ess_matrix(np.ones((3, 3))))                     # test-ESS.R:53-56 (NA + warning)
    assert od.rhat_matrix(rng.standard_normal((1000, 4))) < 1.01        # test-rhat.R:1-5
    drift = np.concatenate([rng.standard_normal(50), rng.standard_normal(50) + 10])[:, None]
    assert od.rhat_matrix(drift) > 2                                    # test-rhat.R:18-27
    assert np.isnan(od.rhat_matrix(np.ones((4, 4))))                    # test-rhat.R:42-45
    with pytest.raises(ValueError, match="Number of iterations must be at least 2"):
        od.rhat_matrix(np.ones((1, 2)))                                 # test-rhat.R:47-50
    assert od.rhat_matrix(rng.standard_normal((1001, 4))) < 1.01        # test-rhat.R:64-69 (odd length)
    near = np.tile(np.array([0.0, 1.0] * 50)[:, None], (1, 2))
    assert od.rhat_matrix(near) == 1.0                                  # R/rhat.R:63-65: [0.99, 1] -> 1


@pytest.fixture(scope="module")
def host_diag(tmp_path_factory):
    exe = tmp_path_factory.mktemp("hd") / "host_diag"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe),
                    os.path.join(ROOT, "tests", "host_diag.cpp")], check=True)

    def run(draws, burn_in=0, want_ess=True):
        k, m, p = draws.shape
        r = subprocess.run([str(exe), str(k), str(m), str(p), str(burn_in), str(int(want_ess))],
                           input=np.ascontiguousarray(draws, dtype=np.float64).tobytes(), capture_output=True)
        assert r.returncode == 0, r.stderr.decode()
        rows = [ln.split() for ln in r.stdout.decode().strip().splitlines()]
        return (np.array([float(x[0]) for x in rows]), np.array([float(x[1]) for x in rows]),
                np.array([int(x[2]) for x in rows]))
    return run


@pytest.mark.parametrize("k,m,p,burn", [(2, 2, 1, 0), (2, 3, 1, 0), (3, 7, 2, 0), (4, 200, 3, 0), (4, 257, 3, 56),
                                        (16, 101, 2, 1), (3, 1000, 1, 0)])
def test_device_thread_bodies_equal_the_oracle(host_diag, k, m, p, burn):
    rng = np.random.default_rng(100 * k + m)
    draws = np.stack([ar1(rng, m, p, rho) * (1 + c) + c for c, rho in zip(range(k), np.linspace(0.0, 0.95, k))], axis=0)
    draws[..., -1] *= 1e-3                                              # parameters on different scales
    ess, rhat, flags = host_diag(draws, burn)
    for j in range(p):
        mat = draws[:, burn:, j].T
        if m - burn >= 4:                                               # m = 2, 3: var of a 1-point half is NA in R
            np.testing.assert_allclose(rhat[j], od.rhat_matrix(mat), rtol=1e-12)
        np.testing.assert_allclose(ess[j], od.ess_matrix(mat), rtol=1e-10)
    assert (flags & 1 == 0).all()


def test_device_thread_bodies_edge_cases(host_diag):
    rng = np.random.default_rng(3)
    draws = rng.standard_normal((3, 40, 3))
    draws[1, :, 1] = 2.5                                                # a chain that never moved (parameter 1)
    draws[2, 20:, 2] = -1.0                                             # a second half that never moved (parameter 2)
    ess, rhat, flags = host_diag(draws)
    assert flags.tolist() == [0, 3, 2]
    assert np.isfinite(ess[0]) and np.isnan(ess[1]) and np.isfinite(ess[2])
    assert np.isfinite(rhat[0]) and np.isnan(rhat[1]) and np.isnan(rhat[2])
    np.testing.assert_allclose(ess[2], od.ess_matrix(draws[:, :, 2].T), rtol=1e-10)
    # one chain: rhat only (R/pmmh.R:580-590)
    one = ar1(rng, 100, 2, 0.5)[None]
    _, rhat1, _ = host_diag(one, want_ess=False)
    np.testing.assert_allclose(rhat1, [od.rhat_matrix(one[0, :, j:j + 1]) for j in range(2)], rtol=1e-12)
    # the [0.99, 1] -> 1 clamp
    near = np.tile(np.array([0.0, 1.0] * 50)[None, :, None], (2, 1, 1))
    assert host_diag(near)[1][0] == 1.0


def test_public_functions_check_their_input_like_the_reference():
    import pandas as pd
    rng = np.random.default_rng(0)
    for fn in (b.ess, b.rhat):
        with pytest.raises(TypeError, match="Input must be a matrix or a data frame with a 'chain' column."):
            fn([1, 2, 3])                                               # test-ESS.R:36-41, test-rhat.R:30-34
        with pytest.raises(ValueError, match="Data frame must contain a 'chain' column."):
            fn(pd.DataFrame({"a": [1, 2, 3], "b": [4, 5, 6]}))          # test-ESS.R:58-64, test-rhat.R:35-39
        with pytest.raises(ValueError, match="Not all chains have the same number of iterations"):
            fn(pd.DataFrame({"chain": [1, 1, 1, 1, 1, 2, 2, 2], "param1": rng.standard_normal(8),
                             "param2": rng.standard_normal(8)}))        # test-ESS.R:66-77, test-rhat.R:52-62
        with pytest.raises(ValueError, match="Number of iterations must be at least 2"):
            fn(rng.standard_normal((1, 3)))
    with pytest.raises(ValueError, match="Number of chains must be at least 2"):
        b.ess(rng.standard_normal((6, 1)))
