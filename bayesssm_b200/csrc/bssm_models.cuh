// bssm_models.cuh -- built-in device models (the "operator" slots of the reference:
// init_fn / transition_fn / log_likelihood_fn / aux_log_likelihood_fn / move_fn,
// R/particle_filter-doc.R:11-18).  A user model compiled through NVRTC follows the
// same contract (struct UserModel).
//
// Contract (all members static, R = float or double):
//   D                       state dimension (<= 4)
//   NTHETA, NCONST          parameters / constants; theta_in = [NTHETA + NCONST] doubles
//   NZ_INIT, NU_INIT        normals / uniforms consumed per particle by init
//   NZ_TRANS, NU_TRANS      ... by one transition
//   NZ_MOVE, NU_MOVE        ... by one RMPF move
//   NPAR                    size of the per-filter derived-parameter block (<= 16)
//   HAS_AUX, HAS_MOVE
//   prepare(theta_in, par)  once per filter (e.g. cache log(sigma))
//   init(x, par, z, u); transition(x, par, t, z, u); loglik(y, x, par, t);
//   aux_loglik(y, x, par, t); move(x, y, par, t, z, u)
// z are N(0,1) draws in R; u are uniforms in (0,1), always double.
// Optional members:
//   DYN_U = true + transition_dyn(x, par, t, DynU& du)   the transition draws a data-dependent number of uniforms: du(k) is its k-th
//                           (Philox mode only; bssm_common.cuh).  transition() is then never called.
//   PACKED = true + transition2(F2 x, par, t, F2 z), loglik2(y, F2 x, par, t)   1-D models: the same operations on a packed pair of
//                           fp32 particles (FADD2 / FMUL2 / FFMA2), used by the streaming engine in the throughput precision.
#pragma once
#include "bssm_common.cuh"

namespace bssm {

// README.md:137-146 (also tests/testthat/test-bootstrap_filter.R:152-164)
struct ModelArSin {
  static constexpr int D = 1, NTHETA = 3, NCONST = 0, NZ_INIT = 1, NU_INIT = 0, NZ_TRANS = 1, NU_TRANS = 0,
                       NZ_MOVE = 1, NU_MOVE = 1, NPAR = 4;
  static constexpr bool HAS_AUX = true, HAS_MOVE = true, COS_OBS = false, HAS_SIN = true;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) {
    par[0] = (R)th[0]; par[1] = (R)th[1]; par[2] = (R)th[2]; par[3] = (R)log(th[2]);
  }
  template <typename R> static BSSM_DEV void init(R* x, const R*, const R* z, const double*) { x[0] = z[0]; }
  // an upper bound of loglik over all states and observations (the persistent kernel's reference for exp(lw - bound))
  template <typename R> static BSSM_DEV R loglik_bound(const R* par) { return -((R)0.918938533204672741780329736406 + par[3]); }
  template <typename R> static BSSM_DEV void transition(R* x, const R* par, int, const R* z, const double*) {
    x[0] = par[0] * x[0] + Math<R>::sin_(x[0]) + par[1] * z[0];
  }
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], x[0], par[2], par[3]);
  }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], par[0] * x[0] + Math<R>::sin_(x[0]), par[2], par[3]);
  }
  // two particles per instruction (throughput precision, streaming engine): transition<float> / loglik<float> lane by lane
  static constexpr bool PACKED = true;
  static BSSM_DEV F2 transition2(F2 x, const float* par, int, F2 z) {
    return f2_fma(f2_make(par[1], par[1]), z, f2_fma(f2_make(par[0], par[0]), x, Math<float>::sin2_(x)));
  }
  static BSSM_DEV F2 loglik2(const double* y, F2 x, const float* par, int) { return dnorm_log2((float)y[0], x, par[2], par[3]); }
  template <typename R> static BSSM_DEV void move(R* x, const double* y, const R* par, int t, const R* z, const double* u) {
    R prop = x[0] + (R)0.1 * z[0];  // tests/testthat/test-resample_move_filter.R:24-35
    R lc = loglik<R>(y, x, par, t), lp = loglik<R>(y, &prop, par, t);
    if (log(u[0]) < (double)(lp - lc)) x[0] = prop;
  }
};

// R/pmmh.R:157-159: same dynamics, y ~ N(cos x, sigma_y)
struct ModelArCos : ModelArSin {
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], Math<R>::cos_(x[0]), par[2], par[3]);
  }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], Math<R>::cos_(par[0] * x[0] + Math<R>::sin_(x[0])), par[2], par[3]);
  }
  static BSSM_DEV F2 loglik2(const double* y, F2 x, const float* par, int) { return dnorm_log2((float)y[0], Math<float>::cos2_(x), par[2], par[3]); }
  template <typename R> static BSSM_DEV void move(R* x, const double* y, const R* par, int t, const R* z, const double* u) {
    R prop = x[0] + (R)0.1 * z[0];
    R lc = loglik<R>(y, x, par, t), lp = loglik<R>(y, &prop, par, t);
    if (log(u[0]) < (double)(lp - lc)) x[0] = prop;
  }
};

// tests/testthat/test-pmmh_tuning.R:163-173 with general (sigma_x, sigma_y)
struct ModelLG {
  static constexpr int D = 1, NTHETA = 3, NCONST = 0, NZ_INIT = 1, NU_INIT = 0, NZ_TRANS = 1, NU_TRANS = 0,
                       NZ_MOVE = 1, NU_MOVE = 1, NPAR = 4;
  static constexpr bool HAS_AUX = true, HAS_MOVE = true;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) {
    par[0] = (R)th[0]; par[1] = (R)th[1]; par[2] = (R)th[2]; par[3] = (R)log(th[2]);
  }
  template <typename R> static BSSM_DEV void init(R* x, const R*, const R* z, const double*) { x[0] = z[0]; }
  template <typename R> static BSSM_DEV R loglik_bound(const R* par) { return -((R)0.918938533204672741780329736406 + par[3]); }
  template <typename R> static BSSM_DEV void transition(R* x, const R* par, int, const R* z, const double*) {
    x[0] = par[0] * x[0] + par[1] * z[0];
  }
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], x[0], par[2], par[3]);
  }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], par[0] * x[0], par[2], par[3]);
  }
  static constexpr bool PACKED = true;
  static BSSM_DEV F2 transition2(F2 x, const float* par, int, F2 z) { return f2_fma(f2_make(par[1], par[1]), z, f2_mul(f2_make(par[0], par[0]), x)); }
  static BSSM_DEV F2 loglik2(const double* y, F2 x, const float* par, int) { return dnorm_log2((float)y[0], x, par[2], par[3]); }
  template <typename R> static BSSM_DEV void move(R* x, const double* y, const R* par, int t, const R* z, const double* u) {
    R prop = x[0] + (R)0.1 * z[0];
    R lc = loglik<R>(y, x, par, t), lp = loglik<R>(y, &prop, par, t);
    if (log(u[0]) < (double)(lp - lc)) x[0] = prop;
  }
};

// tests/testthat/test-auxiliary_filter.R:17-27, test-resample_move_filter.R:17-35
struct ModelRwDrift {
  static constexpr int D = 1, NTHETA = 2, NCONST = 0, NZ_INIT = 1, NU_INIT = 0, NZ_TRANS = 1, NU_TRANS = 0,
                       NZ_MOVE = 1, NU_MOVE = 1, NPAR = 3;
  static constexpr bool HAS_AUX = true, HAS_MOVE = true;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) {
    par[0] = (R)th[0]; par[1] = (R)th[1]; par[2] = (R)log(th[1]);
  }
  template <typename R> static BSSM_DEV void init(R* x, const R*, const R* z, const double*) { x[0] = z[0]; }
  template <typename R> static BSSM_DEV R loglik_bound(const R* par) { return -((R)0.918938533204672741780329736406 + par[2]); }
  template <typename R> static BSSM_DEV void transition(R* x, const R* par, int, const R* z, const double*) {
    x[0] = x[0] + (par[0] + z[0]);
  }
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], x[0], par[1], par[2]);
  }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int) {
    return dnorm_log<R>((R)y[0], x[0] + par[0], par[1], par[2]);
  }
  static constexpr bool PACKED = true;
  static BSSM_DEV F2 transition2(F2 x, const float* par, int, F2 z) { return f2_add(x, f2_add(f2_make(par[0], par[0]), z)); }
  static BSSM_DEV F2 loglik2(const double* y, F2 x, const float* par, int) { return dnorm_log2((float)y[0], x, par[1], par[2]); }
  template <typename R> static BSSM_DEV void move(R* x, const double* y, const R* par, int t, const R* z, const double* u) {
    R prop = x[0] + (R)0.1 * z[0];
    R lc = loglik<R>(y, x, par, t), lp = loglik<R>(y, &prop, par, t);
    if (log(u[0]) < (double)(lp - lc)) x[0] = prop;
  }
};

// Chain-binomial stochastic SIR: state (S, I), daily step
//   new_inf ~ Bin(S, 1 - exp(-lambda I / pop)),  new_rec ~ Bin(I, 1 - exp(-gamma)),
// Poisson observation of I (vignettes/articles/stochastic-sir-model.Rmd:285-310).
struct ModelSirCB {
  static constexpr int D = 2, NTHETA = 2, NCONST = 2, NZ_INIT = 0, NU_INIT = 0, NZ_TRANS = 0, NU_TRANS = 2,
                       NZ_MOVE = 0, NU_MOVE = 2, NPAR = 4;
  static constexpr bool HAS_AUX = true, HAS_MOVE = true;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) {
    par[0] = (R)th[0]; par[1] = (R)th[1]; par[2] = (R)th[2]; par[3] = (R)th[3];
  }
  template <typename R> static BSSM_DEV void init(R* x, const R* par, const R*, const double*) {
    x[0] = par[2] - par[3]; x[1] = par[3];
  }
  template <typename R> static BSSM_DEV void transition(R* x, const R* par, int, const R*, const double* u) {
    if constexpr (sizeof(R) == 4) {    // throughput precision: fp32 inversion on log(1 - p) = -rate (counts up to 2^24 are exact in fp32)
      const float S = x[0], I = x[1];
      if (I == 0.f) return;
      const float ni = binom_inversion_f32(S, -par[0] * I * Math<float>::rcp_(par[2]), (float)u[0]);
      const float nr = binom_inversion_f32(I, -par[1], (float)u[1]);
      x[0] = S - ni; x[1] = I + ni - nr;
      return;
    }
    double S = (double)x[0], I = (double)x[1];
    if (I == 0.0) return;
    double p_inf = 1.0 - exp(-(double)par[0] * I / (double)par[2]);
    double p_rec = 1.0 - exp(-(double)par[1]);
    double ni = binom_inversion(S, p_inf, u[0]);
    double nr = binom_inversion(I, p_rec, u[1]);
    x[0] = (R)(S - ni); x[1] = (R)(I + ni - nr);
  }
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R*, int) {
    return dpois_log<R>((R)y[0], x[1]);
  }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int) {
    if constexpr (sizeof(R) == 4) {
      const float S = x[0], I = x[1];
      const float p_inf = 1.0f - Math<float>::exp_(-par[0] * I * Math<float>::rcp_(par[2])), p_rec = 1.0f - Math<float>::exp_(-par[1]);
      return dpois_log<float>((float)y[0], I + S * p_inf - I * p_rec);
    }
    double S = (double)x[0], I = (double)x[1];
    double p_inf = 1.0 - exp(-(double)par[0] * I / (double)par[2]), p_rec = 1.0 - exp(-(double)par[1]);
    return dpois_log<R>((R)y[0], (R)(I + S * p_inf - I * p_rec));
  }
  template <typename R> static BSSM_DEV void move(R* x, const double* y, const R* par, int t, const R*, const double* u) {
    R prop[2] = {x[0], x[1] + (u[0] < 0.5 ? (R)-1 : (R)1)};
    if (prop[1] < (R)0 || prop[1] > par[2] - x[0]) return;
    R lc = loglik<R>(y, x, par, t), lp = loglik<R>(y, prop, par, t);
    if (log(u[1]) < (double)(lp - lc)) x[1] = prop[1];
  }
};

// Stochastic SIR with the EXACT (Gillespie) daily step of the reference's vignette
// (vignettes/articles/stochastic-sir-model.Rmd:152-176: epidemic_step): competing exponential clocks for infection
// (rate lambda s i / pop) and removal (rate gamma i) until the day is over; two uniforms per event, drawn on demand (DynU).
// Initial state, observation model, auxiliary predictor and move as the chain-binomial model above.
struct ModelSirGillespie : ModelSirCB {
  static constexpr int NU_TRANS = 0;
  static constexpr bool DYN_U = true;
  template <typename R> static BSSM_DEV void transition_dyn(R* x, const R* par, int, DynU& du) {
    // always fp64: the event times of a day sum to ~1 and the comparison t + dt > 1 ends the loop; the work is the
    // data-dependent event count (tens to hundreds of events per particle and day), not the arithmetic
    double s = (double)x[0], i = (double)x[1], t = 0.0;
    const double lam = (double)par[0] / (double)par[2], gam = (double)par[1];
    const int max_events = 2 * (int)par[2] + 8;          // s + 2 i <= 2 pop events can happen at all
    for (int e = 0; e < max_events && i > 0.0; e++) {
      const double rate_inf = lam * s * i, rate_rem = gam * i, rate = rate_inf + rate_rem;
      if (!(rate > 0.0)) break;
      const double dt = -log(du(2 * e)) / rate;           // rexp(1, rate)
      if (t + dt > 1.0) break;
      t += dt;
      if (du(2 * e + 1) < rate_inf / rate) { s -= 1.0; i += 1.0; } else { i -= 1.0; }
    }
    x[0] = (R)s; x[1] = (R)i;
  }
  template <typename R> static BSSM_DEV void transition(R*, const R*, int, const R*, const double*) {}   // (never called: DYN_U)
};

// tests/testthat/test-bootstrap_filter.R:211-217, test-pmmh.R:622-628: 2-D random walk, flat likelihood
struct ModelRw2D {
  static constexpr int D = 2, NTHETA = 1, NCONST = 0, NZ_INIT = 2, NU_INIT = 0, NZ_TRANS = 2, NU_TRANS = 0,
                       NZ_MOVE = 0, NU_MOVE = 0, NPAR = 1;
  static constexpr bool HAS_AUX = false, HAS_MOVE = false;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) { par[0] = (R)th[0]; }
  template <typename R> static BSSM_DEV void init(R* x, const R*, const R* z, const double*) { x[0] = z[0]; x[1] = z[1]; }
  template <typename R> static BSSM_DEV void transition(R* x, const R* par, int, const R* z, const double*) {
    x[0] = x[0] + (par[0] + z[0]); x[1] = x[1] + (par[0] + z[1]);
  }
  template <typename R> static BSSM_DEV R loglik(const double*, const R*, const R*, int) { return (R)1; }
  template <typename R> static BSSM_DEV R aux_loglik(const double*, const R*, const R*, int) { return (R)1; }
  template <typename R> static BSSM_DEV void move(R*, const double*, const R*, int, const R*, const double*) {}
};

}  // namespace bssm
