"""User models as CUDA snippets compiled by NVRTC (the GPU counterpart of passing R closures; precedent:
the cppFunction transition of vignettes/articles/detailed-overview.Rmd:408-466)."""
import ctypes as C

import numpy as np
import pytest

import bayesssm_b200 as b
import engine_helpers as eh
from bayesssm_b200 import _native as nat
from test_filter_gpu import sim_y

pytestmark = pytest.mark.gpu

# the README model (README.md:137-146) restated as a snippet: must reproduce the built-in model bit for bit
AR_SNIPPET = r'''
struct UserModel {
  static constexpr int D = 1, NTHETA = 3, NCONST = 0, NZ_INIT = 1, NU_INIT = 0, NZ_TRANS = 1, NU_TRANS = 0,
                       NZ_MOVE = 1, NU_MOVE = 1, NPAR = 4;
  static constexpr bool HAS_AUX = true, HAS_MOVE = true;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) {
    par[0] = (R)th[0]; par[1] = (R)th[1]; par[2] = (R)th[2]; par[3] = (R)log(th[2]);
  }
  template <typename R> static BSSM_DEV void init(R* x, const R*, const R* z, const double*) { x[0] = z[0]; }
  template <typename R> static BSSM_DEV void transition(R* x, const R* par, int, const R* z, const double*) {
    x[0] = par[0] * x[0] + Math<R>::sin_(x[0]) + par[1] * z[0];
  }
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], x[0], par[2], par[3]);
  }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], par[0] * x[0] + Math<R>::sin_(x[0]), par[2], par[3]);
  }
  template <typename R> static BSSM_DEV void move(R* x, const double* y, const R* par, int t, const R* z, const double* u) {
    R prop = x[0] + (R)0.1 * z[0];
    R lc = loglik<R>(y, x, par, t), lp = loglik<R>(y, &prop, par, t);
    if (log(u[0]) < (double)(lp - lc)) x[0] = prop;
  }
};
'''

# stochastic volatility: a model the engine does not ship
SV_SNIPPET = r'''
struct UserModel {
  static constexpr int D = 1, NTHETA = 3, NCONST = 0, NZ_INIT = 1, NU_INIT = 0, NZ_TRANS = 1, NU_TRANS = 0,
                       NZ_MOVE = 0, NU_MOVE = 0, NPAR = 4;
  static constexpr bool HAS_AUX = false, HAS_MOVE = false;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) {
    par[0] = (R)th[0]; par[1] = (R)th[1]; par[2] = (R)th[2]; par[3] = (R)(th[2] / sqrt(1.0 - th[1] * th[1]));
  }
  template <typename R> static BSSM_DEV void init(R* x, const R* par, const R* z, const double*) { x[0] = par[0] + par[3] * z[0]; }
  template <typename R> static BSSM_DEV void transition(R* x, const R* par, int, const R* z, const double*) {
    x[0] = par[0] + par[1] * (x[0] - par[0]) + par[2] * z[0];
  }
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R*, int) {
    R yy = (R)y[0];
    return -((R)0.918938533204672741780329736406 + (R)0.5 * x[0] + (R)0.5 * yy * yy * Math<R>::exp_(-x[0]));
  }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int t) { return loglik<R>(y, x, par, t); }
  template <typename R> static BSSM_DEV void move(R*, const double*, const R*, int, const R*, const double*) {}
};
'''


def test_snippet_reproduces_builtin_model_and_oracle(orc, engine):
    mid = C.c_int()
    nat.check(engine.lib.bssm_model_compile(engine.handle, AR_SNIPPET.encode(), C.byref(mid)))
    assert mid.value >= 1000
    rng = np.random.default_rng(1)
    y = sim_y(0, 15, rng)
    th = [0.8, 1.0, 0.5]
    for algorithm in (0, 1, 2):
        noise = orc.make_noise(0, 600, 15, 15, rng)
        ref = orc.particle_filter(0, algorithm, 2, 0, 600, y, th, noise=noise, want_ancestors=True)
        usr = eh.filter_run(engine, mid.value, algorithm, 2, 0, 600, y, th, noise=noise, want_ancestors=True)
        blt = eh.filter_run(engine, 0, algorithm, 2, 0, 600, y, th, noise=noise, want_ancestors=True)
        assert usr["loglike"][0] == blt["loglike"][0] and np.array_equal(usr["state_est"], blt["state_est"])
        assert abs(usr["loglike"][0] - ref["loglike"]) <= 1e-10 * abs(ref["loglike"])
        assert np.array_equal(usr["ancestors_history"][0], ref["ancestors_history"])


def test_new_model_through_the_public_api():
    sv = b.models.cuda_model("stochastic_volatility", SV_SNIPPET, ("mu", "phi", "sigma"))
    rng = np.random.default_rng(2)
    x, ys = -1.0, []
    for _ in range(200):
        x = -1.0 + 0.95 * (x + 1.0) + 0.25 * rng.standard_normal()
        ys.append(np.exp(x / 2) * rng.standard_normal())
    y = np.array(ys)
    lls = [b.bootstrap_filter(y, 20000, sv.init_fn, sv.transition_fn, sv.log_likelihood_fn, return_particles=False,
                              seed=s, mu=-1.0, phi=0.95, sigma=0.25)["loglike"] for s in range(4)]
    bad = b.bootstrap_filter(y, 20000, sv.init_fn, sv.transition_fn, sv.log_likelihood_fn, return_particles=False,
                             seed=0, mu=2.0, phi=0.95, sigma=0.25)["loglike"]
    assert np.isfinite(lls).all() and np.std(lls) < 0.5 and bad < min(lls) - 5      # the true mu is far more likely
    pri = {"mu": b.priors.normal(0, 2), "phi": b.priors.uniform(0, 1), "sigma": b.priors.exponential(1)}
    with pytest.warns(UserWarning):
        out = b.pmmh(b.bootstrap_filter, y, 200, sv.init_fn, sv.transition_fn, sv.log_likelihood_fn, pri,
                     [{"mu": -1.0, "phi": 0.9, "sigma": 0.3}] * 2, burn_in=50, num_chains=2,
                     param_transform={"mu": "identity", "phi": "logit", "sigma": "log"},
                     tune_control=b.default_tune_control(pilot_m=100, pilot_reps=8), seed=4, print_result=False)
    assert len(out["theta_chain"]) == 300 and np.isfinite(out["theta_chain"][["mu", "phi", "sigma"]].to_numpy()).all()


def test_compile_error_is_reported(engine):
    mid = C.c_int()
    st = engine.lib.bssm_model_compile(engine.handle, b"struct UserModel { this is not C++ };", C.byref(mid))
    assert st == nat.ERR_NVRTC
    assert "error" in nat.last_error().lower()
    assert b"user_model.cu" in engine.lib.bssm_model_compile_log(engine.handle)


def test_snippets_run_on_the_streaming_engine(orc, engine):
    """The streaming engine's kernels are compiled by NVRTC together with the snippet: the README model as a snippet
    reproduces the built-in model bit for bit on that engine too, a new model agrees with the general kernels, and
    AUTO sends f32 bootstrap filters of an eligible snippet there (2 launches per observation, not ~10)."""
    ST = nat.ENGINE_STREAM
    mid, sv = C.c_int(), C.c_int()
    nat.check(engine.lib.bssm_model_compile(engine.handle, AR_SNIPPET.encode(), C.byref(mid)))
    nat.check(engine.lib.bssm_model_compile(engine.handle, SV_SNIPPET.encode(), C.byref(sv)))
    rng = np.random.default_rng(3)
    y = sim_y(0, 25, rng)
    th = [0.8, 1.0, 0.5]
    for prec in (nat.F64, nat.F32):
        for N in (3000, 70001):
            usr = eh.filter_run(engine, mid.value, 0, 2, 0, N, y, th, seed=5, stream_base=2, precision=prec, engine=ST)
            blt = eh.filter_run(engine, 0, 0, 2, 0, N, y, th, seed=5, stream_base=2, precision=prec, engine=ST)
            assert usr["status"][0] == 0
            assert usr["loglike"][0] == blt["loglike"][0] and np.array_equal(usr["state_est"], blt["state_est"])
            assert np.array_equal(usr["ess"], blt["ess"])
    ref = orc.particle_filter(0, 0, 2, 0, 3000, y, th, seed=5, stream=2)
    usr = eh.filter_run(engine, mid.value, 0, 2, 0, 3000, y, th, seed=5, stream_base=2, precision=nat.F64, engine=ST)
    assert abs(usr["loglike"][0] - ref["loglike"]) <= 1e-6 * abs(ref["loglike"])
    # stochastic volatility: streaming engine against the general kernels, same Philox streams
    ysv = 0.6 * rng.standard_normal(40)
    thv = [-1.0, 0.95, 0.25]
    a = eh.filter_run(engine, sv.value, 0, 2, 1, 50000, ysv, thv, seed=8, precision=nat.F64, engine=ST)
    g = eh.filter_run(engine, sv.value, 0, 2, 1, 50000, ysv, thv, seed=8, precision=nat.F64, engine=nat.ENGINE_GENERAL)
    assert a["n_resampled"][0] == g["n_resampled"][0]
    assert abs(a["loglike"][0] - g["loglike"][0]) <= 1e-9 * abs(g["loglike"][0])
    np.testing.assert_allclose(a["state_est"][0], g["state_est"][0], rtol=1e-8, atol=1e-9)
    n0 = engine.launch_count()
    f = eh.filter_run(engine, sv.value, 0, 2, 0, 1 << 18, ysv, thv, seed=8, precision=nat.F32)   # AUTO
    assert engine.launch_count() - n0 < 2 * 40 + 12
    assert abs(f["loglike"][0] - g["loglike"][0]) < 0.3
    # APF / injected noise / 2-D models stay on the general kernels; asking for the streaming engine there is an error
    with pytest.raises(nat.EngineError):
        eh.filter_run(engine, sv.value, 1, 2, 0, 1000, ysv, thv, seed=8, precision=nat.F32, engine=ST)


# a user model that draws its transition uniforms on demand (DYN_U): the exact-Gillespie SIR step of the reference's vignette
# (vignettes/articles/stochastic-sir-model.Rmd:152-176) restated as a snippet -- must reproduce the built-in model 6 bit for bit
GILLESPIE_SNIPPET = r'''
struct UserModel {
  static constexpr int D = 2, NTHETA = 2, NCONST = 2, NZ_INIT = 0, NU_INIT = 0, NZ_TRANS = 0, NU_TRANS = 0,
                       NZ_MOVE = 0, NU_MOVE = 0, NPAR = 4;
  static constexpr bool HAS_AUX = false, HAS_MOVE = false, DYN_U = true;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) { for (int k = 0; k < 4; k++) par[k] = (R)th[k]; }
  template <typename R> static BSSM_DEV void init(R* x, const R* par, const R*, const double*) { x[0] = par[2] - par[3]; x[1] = par[3]; }
  template <typename R> static BSSM_DEV void transition(R*, const R*, int, const R*, const double*) {}
  template <typename R> static BSSM_DEV void transition_dyn(R* x, const R* par, int, DynU& du) {
    double s = (double)x[0], i = (double)x[1], t = 0.0;
    const double lam = (double)par[0] / (double)par[2], gam = (double)par[1];
    const int max_events = 2 * (int)par[2] + 8;
    for (int e = 0; e < max_events && i > 0.0; e++) {
      const double rate_inf = lam * s * i, rate_rem = gam * i, rate = rate_inf + rate_rem;
      if (!(rate > 0.0)) break;
      const double dt = -log(du(2 * e)) / rate;
      if (t + dt > 1.0) break;
      t += dt;
      if (du(2 * e + 1) < rate_inf / rate) { s -= 1.0; i += 1.0; } else { i -= 1.0; }
    }
    x[0] = (R)s; x[1] = (R)i;
  }
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R*, int) { return dpois_log<R>((R)y[0], x[1]); }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int t) { return loglik<R>(y, x, par, t); }
  template <typename R> static BSSM_DEV void move(R*, const double*, const R*, int, const R*, const double*) {}
};
'''


def test_user_model_with_uniforms_on_demand(orc, engine):
    mid = C.c_int()
    nat.check(engine.lib.bssm_model_compile(engine.handle, GILLESPIE_SNIPPET.encode(), C.byref(mid)))
    y = np.array([82, 95, 118, 130, 151, 160], dtype=float)
    th = [0.5, 0.2, 500.0, 70.0]
    for prec in (nat.F64, nat.F32):
        a = eh.filter_run(engine, mid.value, 0, 2, 0, 3000, y, th, seed=5, run_id=2, stream_base=1, precision=prec, num_filters=2)
        b_ = eh.filter_run(engine, 6, 0, 2, 0, 3000, y, th, seed=5, run_id=2, stream_base=1, precision=prec, num_filters=2)
        np.testing.assert_array_equal(a["loglike"], b_["loglike"])
        np.testing.assert_array_equal(a["state_est"], b_["state_est"])
    ref = orc.particle_filter(6, 0, 2, 0, 3000, y, th, seed=5, run_id=2, stream=2)
    assert abs(a["loglike"][1] - ref["loglike"]) < 0.05       # (f32 weights; the trajectories are the oracle's)
    # injected noise buffers cannot serve it
    noise = orc.make_noise(3, 64, 2, 2, np.random.default_rng(0))
    with pytest.raises(Exception, match="on demand"):
        eh.filter_run(engine, mid.value, 0, 2, 0, 64, y[:2], th, noise=noise)
