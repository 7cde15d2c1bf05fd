// bssm_pmmh.cuh -- device side of the Particle Marginal Metropolis-Hastings driver (bssm_pmmh.cu): the per-chain
// state, the R densities / transforms (SURVEY.md Appendix F, R/utils.R:102-152) and the one-thread-per-chain kernels
// (start, first draw, proposal, accept / reject, latent state bookkeeping, pilot statistics, replicate set-up, tuning).
// A header so that the same text is compiled by nvcc for the library and by g++ for the CPU logic test
// (tests/host_pmmh.cpp over tests/simt_emu.h).
#pragma once
#include "../../include/bayesssm_b200.h"
#include "bssm_common.cuh"

#include <math.h>

namespace bssm {


enum { PH_PILOT = 1, PH_PILOT_RUN = 2, PH_MAIN = 3 };
constexpr int PMAX = 8;

struct PmmhDev {
  int C, p, nconst, theta_stride;
  unsigned long long seed;
  unsigned int chain_id_base;
  int prior_kind[PMAX]; double prior_a[PMAX], prior_b[PMAX];
  int transform[PMAX]; double pilot_sd[PMAX];
  double consts[8];
  // per chain
  double *cur, *prop;        // [C][p]
  double *cur_ll;            // [C]
  double *lp_prop;           // [C] sum of log priors at the proposal
  double *theta_full;        // [C][theta_stride]  filter input (theta, consts)
  int *valid;                // [C] proposal has finite priors (main chain) / chain alive
  int *alive;                // [C]
  int *status;               // [C]
  int *n_accept;             // [C]
  int *moved;                // [C] 1 if the chain took a new state in this iteration (latent state bookkeeping)
  unsigned int *stream, *run_id;  // [C] filter Philox ids
  // filter outputs
  const double* f_loglike; const int* f_status;
  const int* f_nres;         // [C] resampling steps of the last filter run
  unsigned long long* nres_acc;  // [2]: resampling steps / filter runs of the main phase, summed over the chains (for the roofline)
};

// R densities as in SURVEY.md Appendix F
__device__ inline double dev_log_prior(int kind, double a, double b, double x) {
  const double LSP = 0.918938533204672741780329736406;
  const double INF = __longlong_as_double(0x7FF0000000000000LL);
  switch (kind) {
    case BSSM_PRIOR_FLAT: return 0.0;
    case BSSM_PRIOR_NORMAL: { double z = (x - a) / b; return -(LSP + 0.5 * z * z + log(b)); }
    case BSSM_PRIOR_EXP: return x < 0 ? -INF : log(a) - a * x;
    case BSSM_PRIOR_UNIF: return (a <= x && x <= b) ? -log(b - a) : -INF;
    case BSSM_PRIOR_HALFNORMAL: { if (x < 0) return -INF; double z = (x - 0.0) / a; return log(2.0) + (-(LSP + 0.5 * z * z + log(a))); }
  }
  return __longlong_as_double(0x7FF8000000000000LL);
}
__device__ inline double dev_transform(double th, int tr) {  // R/utils.R:102-112
  return tr == BSSM_TR_LOG ? log(th) : (tr == BSSM_TR_LOGIT ? log(th / (1.0 - th)) : th);
}
__device__ inline double dev_back_transform(double z, int tr) {  // R/utils.R:122-132
  return tr == BSSM_TR_LOG ? exp(z) : (tr == BSSM_TR_LOGIT ? 1.0 / (1.0 + exp(-z)) : z);
}
__device__ inline double dev_log_jacobian(const double* th, const int* tr, int p) {  // R/utils.R:142-152 (sic, A14)
  double s = 0.0;
  for (int j = 0; j < p; j++) {
    if (tr[j] == BSSM_TR_LOG) s += log(th[j]);
    else if (tr[j] == BSSM_TR_LOGIT) s += log(1.0 / (th[j] * (1.0 - th[j])));
  }
  return s;
}
__device__ inline bool dev_priors_finite(const PmmhDev& P, const double* th, double* sum_out) {
  double s = 0.0; bool ok = true;
  for (int j = 0; j < P.p; j++) {
    double lp = dev_log_prior(P.prior_kind[j], P.prior_a[j], P.prior_b[j], th[j]);
    if (!isfinite(lp)) ok = false;
    s += lp;
  }
  *sum_out = s;
  return ok;
}
__device__ inline double theta_normal(const PmmhDev& P, int phase, unsigned int chain, unsigned int it, unsigned int attempt, int j) {
  NoiseKey key = make_key(P.seed, (unsigned int)phase << 28, chain);
  uint4x q = noise_quad(key, it, TAG_THETA_Z, attempt, (unsigned int)j >> 2);
  int pr = (j & 3) >> 1;
  double n0, n1;
  Math<double>::box_muller(q.w[2 * pr], q.w[2 * pr + 1], n0, n1);
  return (j & 1) ? n1 : n0;
}
__device__ inline double theta_uniform(const PmmhDev& P, int phase, unsigned int chain, unsigned int it) {
  NoiseKey key = make_key(P.seed, (unsigned int)phase << 28, chain);
  uint4x q = noise_quad(key, it, TAG_THETA_U, 0u, 0u);
  return word_to_unit_f64(q.w[0]);
}
__device__ inline void set_filter_theta(const PmmhDev& P, int c, const double* th) {
  double* tf = P.theta_full + (size_t)c * P.theta_stride;
  for (int j = 0; j < P.p; j++) tf[j] = th[j];
  for (int j = 0; j < P.nconst; j++) tf[P.p + j] = P.consts[j];
}

// start of a phase: current point = start[c], valid iff priors finite (R/pmmh_tuning.R:135-143)
__global__ void k_pm_start(PmmhDev P, const double* start, int phase, int check_prior, int init_state) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  double th[PMAX];
  for (int j = 0; j < P.p; j++) th[j] = start[(size_t)c * P.p + j];
  double s;
  bool ok = dev_priors_finite(P, th, &s);
  if (init_state) { P.alive[c] = 1; P.status[c] = 0; P.n_accept[c] = 0; }
  if (check_prior && !ok && P.alive[c]) { P.alive[c] = 0; P.status[c] = BSSM_ERR_PRIOR_INIT; }
  for (int j = 0; j < P.p; j++) { P.cur[(size_t)c * P.p + j] = th[j]; P.prop[(size_t)c * P.p + j] = th[j]; }
  set_filter_theta(P, c, th);
  P.valid[c] = P.alive[c];
  P.stream[c] = P.chain_id_base + (unsigned int)c;
  P.run_id[c] = ((unsigned int)phase << 28) | 0u;
}
// after the first filter of a phase: record its log-likelihood and draw 0
__global__ void k_pm_first(PmmhDev P, double* chain, double* ll_chain, int m) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  if (P.alive[c] && P.f_status[c]) { P.alive[c] = 0; P.status[c] = P.f_status[c]; }
  P.cur_ll[c] = P.f_loglike[c];
  P.moved[c] = 1;
  for (int j = 0; j < P.p; j++) chain[((size_t)c * m) * P.p + j] = P.cur[(size_t)c * P.p + j];
  ll_chain[(size_t)c * m] = P.cur_ll[c];
}

// proposal.  pilot (R/pmmh_tuning.R:193-208): z* = z + N(0, diag(sd^2)), re-drawn until every prior is finite.
// main (R/pmmh.R:424-442): z* ~ N(z, L L'); a non-finite prior rejects WITHOUT running the filter.
__global__ void k_pm_propose(PmmhDev P, int phase, int it, const double* chol /* [C][p][p] */) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  P.run_id[c] = ((unsigned int)phase << 28) | (unsigned int)it;
  if (!P.alive[c]) { P.valid[c] = 0; return; }
  unsigned int chain = P.chain_id_base + (unsigned int)c;
  const int p = P.p;
  double cur[PMAX], prop[PMAX], zc[PMAX];
  for (int j = 0; j < p; j++) { cur[j] = P.cur[(size_t)c * p + j]; zc[j] = dev_transform(cur[j], P.transform[j]); }
  double lp = 0.0;
  bool ok = false;
  if (phase == PH_PILOT) {
    for (unsigned int attempt = 0; attempt <= 0xFFFFu; attempt++) {
      for (int j = 0; j < p; j++) {
        double zp = zc[j] + (0.0 + P.pilot_sd[j] * theta_normal(P, phase, chain, (unsigned int)it, attempt, j));
        prop[j] = dev_back_transform(zp, P.transform[j]);
      }
      if (dev_priors_finite(P, prop, &lp)) { ok = true; break; }
    }
    if (!ok) { P.alive[c] = 0; P.status[c] = BSSM_ERR_BAD_ARG; }
  } else {
    double xi[PMAX];
    for (int j = 0; j < p; j++) xi[j] = theta_normal(P, phase, chain, (unsigned int)it, 0u, j);
    const double* L = chol + (size_t)c * p * p;
    for (int a = 0; a < p; a++) {
      double s = zc[a];
      for (int b = 0; b <= a; b++) s += L[a * p + b] * xi[b];
      prop[a] = dev_back_transform(s, P.transform[a]);
    }
    ok = dev_priors_finite(P, prop, &lp);
  }
  for (int j = 0; j < p; j++) P.prop[(size_t)c * p + j] = prop[j];
  P.lp_prop[c] = lp;
  P.valid[c] = ok ? 1 : 0;
  if (ok) set_filter_theta(P, c, prop);
}

// accept / reject (R/pmmh_tuning.R:233-253, R/pmmh.R:461-496) and draw storage
__global__ void k_pm_accept(PmmhDev P, int phase, int it, double* chain, double* ll_chain, int m) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  const int p = P.p;
  double cur[PMAX];
  for (int j = 0; j < p; j++) cur[j] = P.cur[(size_t)c * p + j];
  P.moved[c] = 0;
  if (P.alive[c] && P.valid[c]) {
    if (phase == PH_MAIN && P.nres_acc) { atomicAdd(&P.nres_acc[0], (unsigned long long)P.f_nres[c]); atomicAdd(&P.nres_acc[1], 1ull); }
    if (P.f_status[c]) { P.alive[c] = 0; P.status[c] = P.f_status[c]; }
    else {
      double prop[PMAX];
      for (int j = 0; j < p; j++) prop[j] = P.prop[(size_t)c * p + j];
      double prop_ll = P.f_loglike[c], cur_ll = P.cur_ll[c];
      double lp_cur;
      dev_priors_finite(P, cur, &lp_cur);
      double jp = dev_log_jacobian(prop, P.transform, p), jc = dev_log_jacobian(cur, P.transform, p);
      double num, den;
      if (phase == PH_PILOT) { num = P.lp_prop[c] + prop_ll + jp; den = lp_cur + cur_ll + jc; }
      else { num = prop_ll + P.lp_prop[c] + jp; den = cur_ll + lp_cur + jc; }
      double ratio = num - den;
      if (ratio != ratio) ratio = -__longlong_as_double(0x7FF0000000000000LL);
      double u = theta_uniform(P, phase, P.chain_id_base + (unsigned int)c, (unsigned int)it);
      if (log(u) < ratio) {
        for (int j = 0; j < p; j++) { cur[j] = prop[j]; P.cur[(size_t)c * p + j] = prop[j]; }
        P.cur_ll[c] = prop_ll;
        P.moved[c] = 1;
        if (phase == PH_MAIN) P.n_accept[c] += 1;
      }
    }
  }
  for (int j = 0; j < p; j++) chain[((size_t)c * m + it) * p + j] = cur[j];
  ll_chain[(size_t)c * m + it] = P.cur_ll[c];
}

// latent state estimates of the chain (R/pmmh.R:420,494-499): the state_est of the filter run that produced the
// current state is carried along and stored for every iteration
__global__ void k_pm_latent(PmmhDev P, const double* f_state_est, int len /* (T+1) d */, double* cur_se, double* se_chain, int it, int m) {
  const int c = blockIdx.y;
  const int moved = P.moved[c];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < len; k += gridDim.x * blockDim.x) {
    double v = moved ? f_state_est[(size_t)c * len + k] : cur_se[(size_t)c * len + k];
    if (moved) cur_se[(size_t)c * len + k] = v;
    se_chain[((size_t)c * m + it) * len + k] = v;
  }
}

// pilot posterior mean / covariance of the second half on the original scale (R/pmmh_tuning.R:260-267)
__global__ void k_pm_pilot_stats(PmmhDev P, const double* chain, int pilot_m, double* mean, double* cov) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  const int p = P.p;
  int b0 = pilot_m / 2, nn = pilot_m - b0;
  const double* ch = chain + (size_t)c * pilot_m * p;
  double mu[PMAX];
  for (int j = 0; j < p; j++) {
    double s = 0.0;
    for (int it = b0; it < pilot_m; it++) s += ch[(size_t)it * p + j];
    mu[j] = s / (double)nn;
    mean[(size_t)c * p + j] = mu[j];
  }
  for (int a = 0; a < p; a++)
    for (int b = 0; b < p; b++) {
      double s = 0.0;
      for (int it = b0; it < pilot_m; it++) s += (ch[(size_t)it * p + a] - mu[a]) * (ch[(size_t)it * p + b] - mu[b]);
      cov[((size_t)c * p + a) * p + b] = s / (double)(nn - 1);
    }
}

// replicate filters of .pilot_run: filter r of chain c is batch entry c*reps + r
__global__ void k_pm_reps_setup(PmmhDev P, const double* mean, int reps, double* theta_rep, unsigned int* stream_rep,
                                unsigned int* run_rep, int* active_rep) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.C * reps) return;
  int c = i / reps, r = i % reps;
  double* tf = theta_rep + (size_t)i * P.theta_stride;
  for (int j = 0; j < P.p; j++) tf[j] = mean[(size_t)c * P.p + j];
  for (int j = 0; j < P.nconst; j++) tf[P.p + j] = P.consts[j];
  stream_rep[i] = P.chain_id_base + (unsigned int)c;
  run_rep[i] = ((unsigned int)PH_PILOT_RUN << 28) | (unsigned int)r;
  active_rep[i] = P.alive[c];
}

// target_n (R/pmmh_tuning.R:54-57) and the proposal factor: lower Cholesky of D Sigma D (R/pmmh.R:378-389)
__global__ void k_pm_tune(PmmhDev P, const double* rep_ll, const int* rep_status, int reps, int pilot_n, int fixed_n,
                          const double* mean, const double* cov, double* pilot_ll_out, int* target_n, double* chol) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  const int p = P.p;
  double mu = 0.0;
  for (int r = 0; r < reps; r++) {
    double v = rep_ll[(size_t)c * reps + r];
    if (pilot_ll_out) pilot_ll_out[(size_t)c * reps + r] = v;
    if (P.alive[c] && rep_status[(size_t)c * reps + r]) { P.alive[c] = 0; P.status[c] = rep_status[(size_t)c * reps + r]; }
    mu += v;
  }
  mu /= (double)reps;
  double v = 0.0;
  for (int r = 0; r < reps; r++) { double dlt = rep_ll[(size_t)c * reps + r] - mu; v += dlt * dlt; }
  v /= (double)(reps - 1);
  double tn = ceil((double)pilot_n * v);
  if (!(tn >= 50)) tn = 50;
  if (tn > 1000) tn = 1000;
  target_n[c] = fixed_n > 0 ? fixed_n : (int)tn;
  double sc[PMAX], S[PMAX * PMAX];
  for (int j = 0; j < p; j++) {
    double th = mean[(size_t)c * p + j];
    sc[j] = P.transform[j] == BSSM_TR_LOG ? 1.0 / th : (P.transform[j] == BSSM_TR_LOGIT ? 1.0 / (th * (1.0 - th)) : 1.0);
  }
  for (int a = 0; a < p; a++) for (int b = 0; b < p; b++) S[a * p + b] = sc[a] * cov[((size_t)c * p + a) * p + b] * sc[b];
  double* L = chol + (size_t)c * p * p;
  for (int i = 0; i < p * p; i++) L[i] = 0.0;
  for (int j = 0; j < p; j++) {
    double dsum = S[j * p + j];
    for (int k = 0; k < j; k++) dsum -= L[j * p + k] * L[j * p + k];
    if (!(dsum > 0)) continue;
    double dj = sqrt(dsum);
    L[j * p + j] = dj;
    for (int i = j + 1; i < p; i++) {
      double s = S[i * p + j];
      for (int k = 0; k < j; k++) s -= L[i * p + k] * L[j * p + k];
      L[i * p + j] = s / dj;
    }
  }
}

}  // namespace bssm
