"""GPU parity of the device-resident PMMH (through the C ABI) against the oracle restating
R/pmmh.R:345-505 and R/pmmh_tuning.R:29-64,111-317, chain by chain, on the same Philox streams."""
import numpy as np
import pytest

import engine_helpers as eh
from bayesssm_b200 import _native as nat

pytestmark = pytest.mark.gpu

# README.md:153-168 priors: phi ~ U(0,1), sigma_x ~ Exp(1), sigma_y ~ Exp(1); transforms logit/log/log
PRIOR = dict(prior_kind=[3, 2, 2], prior_a=[0.0, 1.0, 1.0], prior_b=[1.0, 0.0, 0.0])


def readme_data(T, rng):
    x, ys = rng.standard_normal(), []
    for _ in range(T):
        x = 0.8 * x + np.sin(x) + rng.standard_normal()
        ys.append(x + 0.5 * rng.standard_normal())
    return np.array(ys)


# the last case: tune_control's pilot_resample_algorithm = "SISR", pilot_resample_fn = "systematic" -- they steer the pilot chain AND
# the .pilot_run replicates behind target_n (R/pmmh.R:366-367 -> `...` of .run_pilot_chain -> do.call(.pilot_run), R/pmmh_tuning.R:292-305)
@pytest.mark.parametrize("transform,pilot", [([0, 0, 0], (2, 0)), ([2, 1, 1], (2, 0)), ([2, 1, 1], (1, 1))])
def test_chains_match_oracle(orc, engine, transform, pilot):
    rng = np.random.default_rng(1405)
    y = readme_data(12, rng)
    inits = np.array([[0.8, 1.0, 0.5], [0.5, 0.7, 1.2], [0.3, 1.5, 0.8]])
    kw = dict(transform=transform, pilot_proposal_sd=[0.1, 0.15, 0.2], pilot_n=64, pilot_m=30, pilot_reps=6, m=40, seed=99,
              pilot_resample_algorithm=pilot[0], pilot_resample_fn=pilot[1])
    got = eh.pmmh_run(engine, 0, 0, y, inits, chain_id_base=4, return_latent_state_est=True, **PRIOR, **kw)
    assert (got["status"] == 0).all()
    for c in range(3):
        ref = orc.pmmh_chain(0, 0, y, inits[c], chain_id=4 + c, return_latent_state_est=True, **PRIOR, **kw)
        assert ref["status"] == 0
        np.testing.assert_allclose(got["pilot_theta_chain"][c], ref["pilot_theta_chain"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got["pilot_loglike_chain"][c], ref["pilot_loglike_chain"], rtol=1e-9)
        np.testing.assert_allclose(got["pilot_theta_mean"][c], ref["pilot_theta_mean"], rtol=1e-9)
        np.testing.assert_allclose(got["pilot_theta_cov"][c], ref["pilot_theta_cov"], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(got["pilot_loglikes"][c], ref["pilot_loglikes"], rtol=1e-9)
        assert got["target_n"][c] == ref["target_n"]
        np.testing.assert_allclose(got["proposal_chol"][c], ref["proposal_chol"], rtol=1e-7, atol=1e-12)
        np.testing.assert_allclose(got["theta_chain"][c], ref["theta_chain"], rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(got["loglike_chain"][c], ref["loglike_chain"], rtol=1e-8)
        assert got["n_accept"][c] == ref["n_accept"]
        # R/pmmh.R:420,494-499: the state estimate behind every draw, carried over on rejections
        np.testing.assert_allclose(got["latent_state_chain"][c], ref["latent_state_chain"], rtol=1e-8, atol=1e-10)


def test_invalid_initial_parameters(engine):
    # R/pmmh_tuning.R:135-143: "Initial parameter values are invalid" -> per-chain status
    rng = np.random.default_rng(1)
    y = readme_data(8, rng)
    inits = np.array([[0.8, 1.0, 0.5], [1.5, 1.0, 0.5]])   # phi = 1.5 outside U(0,1)
    got = eh.pmmh_run(engine, 0, 0, y, inits, transform=[0, 0, 0], pilot_proposal_sd=[0.1] * 3, pilot_n=32,
                      pilot_m=10, pilot_reps=4, m=10, seed=1, **PRIOR)
    assert got["status"][0] == 0 and got["status"][1] == nat.ERR_PRIOR_INIT


def test_fixed_particles_and_chain_sharding(engine):
    # chains are keyed by their GLOBAL id: running [0..3] in one call == running [0,1] and [2,3] in two calls
    rng = np.random.default_rng(2)
    y = readme_data(10, rng)
    inits = np.tile([0.7, 1.0, 0.6], (4, 1))
    kw = dict(transform=[2, 1, 1], pilot_proposal_sd=[0.1] * 3, pilot_n=48, pilot_m=16, pilot_reps=4, m=16, seed=5,
              fixed_num_particles=256, **PRIOR)
    full = eh.pmmh_run(engine, 0, 0, y, inits, **kw)
    lo = eh.pmmh_run(engine, 0, 0, y, inits[:2], chain_id_base=0, **kw)
    hi = eh.pmmh_run(engine, 0, 0, y, inits[2:], chain_id_base=2, **kw)
    assert (full["target_n"] == 256).all()
    np.testing.assert_array_equal(full["theta_chain"][:2], lo["theta_chain"])
    np.testing.assert_array_equal(full["theta_chain"][2:], hi["theta_chain"])


def test_flat_likelihood_recovers_prior(engine):
    # tests/testthat/test-pmmh.R:619-668: flat likelihood => posterior of phi = its N(0,1) prior (mean 0 +- 0.1)
    y = np.zeros((10, 1))
    inits = np.zeros((64, 1))
    got = eh.pmmh_run(engine, 5, 0, y, inits, prior_kind=[1], prior_a=[0.0], prior_b=[1.0], transform=[0],
                      pilot_proposal_sd=[1.0], pilot_n=20, pilot_m=200, pilot_reps=4, m=400, seed=3)
    draws = got["theta_chain"][:, 100:, 0]
    assert abs(draws.mean()) < 0.1
    assert abs(draws.std() - 1.0) < 0.15
