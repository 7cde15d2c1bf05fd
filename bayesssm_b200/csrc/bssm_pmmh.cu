// bssm_pmmh.cu -- device-resident Particle Marginal Metropolis-Hastings.
// Replaces the per-chain closure chain_result (R/pmmh.R:345-505), the pilot chain
// .run_pilot_chain (R/pmmh_tuning.R:111-317), .pilot_run (R/pmmh_tuning.R:29-64), the parameter
// transforms (R/utils.R:102-152) and the chain fan-out (R/pmmh.R:511-535).
// All chains of this call advance together: one thread per chain proposes / accepts, the
// batched filter ([chains x particles]) runs in between, and nothing returns to the host until
// a phase is over.  Quirks A10-A16 of SURVEY.md Appendix A are reproduced on purpose.
#include "bssm_engine.cuh"
#include "bssm_pmmh.cuh"

#include <math.h>

using namespace bssm;

namespace bssm {

struct PmmhBuffers {
  double *cur, *prop, *cur_ll, *lp_prop, *theta_full, *pilot_chain, *pilot_ll, *mean, *cov, *chol, *chain, *ll_chain,
      *theta_rep, *y;
  int *valid, *alive, *status, *n_accept, *target_n, *active_rep, *obs;
  unsigned int *ids, *ids_rep;
};

// one batched filter pass for the chains (theta from P.theta_full, activity from P.valid)
static int run_chain_filters(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, double* cdf, const int* d_active) {
  BSSM_TRY(filter_reset(ctx, f, d_active));
  return filter_enqueue(ctx, f, L, cdf);
}

}  // namespace bssm

extern "C" {

int bssm_pmmh_run(bssm_ctx* ctx, const bssm_pmmh_config* cfg, const double* y, const double* init_theta,
                  bssm_pmmh_result* res) {
  if (!ctx || !cfg || !y || !init_theta || !res) { set_error("bssm_pmmh_run: null argument"); return BSSM_ERR_BAD_ARG; }
  int d, nth, nc;
  BSSM_TRY(model_dims(ctx, cfg->model, &d, &nth, &nc));
  const int C = cfg->num_chains, p = cfg->p, T = cfg->num_obs;
  if (p != nth || p < 1 || p > PMAX) { set_error("pmmh: p=%d does not match the model's %d parameters", p, nth); return BSSM_ERR_BAD_ARG; }
  if (cfg->nconst != nc) { set_error("pmmh: model needs %d constants, got %d", nc, cfg->nconst); return BSSM_ERR_BAD_ARG; }
  if (C < 1 || cfg->m < 1 || T < 1) { set_error("pmmh: num_chains, m and num_obs must be >= 1"); return BSSM_ERR_BAD_ARG; }
  if (!cfg->skip_pilot && (cfg->pilot_m < 4 || cfg->pilot_n < 1 || cfg->pilot_reps < 2)) { set_error("pmmh: pilot_m >= 4, pilot_n >= 1, pilot_reps >= 2 required"); return BSSM_ERR_BAD_ARG; }
  if (cfg->skip_pilot && (!cfg->proposal_chol_in || cfg->fixed_num_particles < 1)) { set_error("pmmh: skip_pilot needs proposal_chol_in and fixed_num_particles"); return BSSM_ERR_BAD_ARG; }
  BSSM_CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int m = cfg->m, pm = cfg->skip_pilot ? 1 : cfg->pilot_m, reps = cfg->skip_pilot ? 1 : cfg->pilot_reps;
  const int ts = nth + nc;

  // ---- device state ----
  PmmhBuffers B;
  int slot = SL_P_BASE;
  auto dalloc = [&](double** ptr, size_t n) { return scratch(ctx, slot++, n, ptr); };
  BSSM_TRY(dalloc(&B.cur, (size_t)C * p)); BSSM_TRY(dalloc(&B.prop, (size_t)C * p));
  BSSM_TRY(dalloc(&B.cur_ll, (size_t)C)); BSSM_TRY(dalloc(&B.lp_prop, (size_t)C));
  BSSM_TRY(dalloc(&B.theta_full, (size_t)C * ts));
  BSSM_TRY(dalloc(&B.pilot_chain, (size_t)C * pm * p)); BSSM_TRY(dalloc(&B.pilot_ll, (size_t)C * pm));
  BSSM_TRY(dalloc(&B.mean, (size_t)C * p)); BSSM_TRY(dalloc(&B.cov, (size_t)C * p * p)); BSSM_TRY(dalloc(&B.chol, (size_t)C * p * p));
  BSSM_TRY(dalloc(&B.chain, (size_t)C * m * p)); BSSM_TRY(dalloc(&B.ll_chain, (size_t)C * m));
  BSSM_TRY(dalloc(&B.theta_rep, (size_t)C * reps * ts));
  BSSM_TRY(dalloc(&B.y, (size_t)T * cfg->dy));
  int* ibuf;
  BSSM_TRY(scratch(ctx, slot++, (size_t)C * 5 + (size_t)C * reps + T, &ibuf));
  B.valid = ibuf; B.alive = ibuf + C; B.status = ibuf + 2 * C; B.n_accept = ibuf + 3 * C; B.target_n = ibuf + 4 * C;
  B.active_rep = ibuf + 5 * C; B.obs = B.active_rep + (size_t)C * reps;
  BSSM_TRY(scratch(ctx, slot++, (size_t)2 * C, &B.ids));
  BSSM_TRY(scratch(ctx, slot++, (size_t)2 * C * reps, &B.ids_rep));
  double* d_init;
  BSSM_TRY(scratch(ctx, slot++, (size_t)C * p, &d_init));
  int* d_moved;
  BSSM_TRY(scratch(ctx, slot++, (size_t)C, &d_moved));
  unsigned long long* d_nres;
  BSSM_TRY(scratch(ctx, slot++, (size_t)2, &d_nres));
  BSSM_CK(cudaMemsetAsync(d_nres, 0, sizeof(unsigned long long) * 2, st));
  // latent state estimates (R/pmmh.R: return_latent_state_est): current + per-iteration copies, only when asked for
  const bool latent = cfg->return_latent_state_est != 0 && res->latent_state_chain != nullptr;
  const int se_len = (T + 1) * d;
  double *d_cur_se = nullptr, *d_se_chain = nullptr;
  if (latent) {
    BSSM_TRY(scratch(ctx, slot++, (size_t)C * se_len, &d_cur_se));
    BSSM_TRY(scratch(ctx, slot++, (size_t)C * m * se_len, &d_se_chain));
  }
  BSSM_CK(cudaMemcpyAsync(B.y, y, sizeof(double) * T * cfg->dy, cudaMemcpyHostToDevice, st));
  BSSM_CK(cudaMemcpyAsync(d_init, init_theta, sizeof(double) * C * p, cudaMemcpyHostToDevice, st));
  if (cfg->obs_times) BSSM_CK(cudaMemcpyAsync(B.obs, cfg->obs_times, sizeof(int) * T, cudaMemcpyHostToDevice, st));
  BSSM_CK(cudaMemsetAsync(B.alive, 0, sizeof(int) * C * 5, st));

  PmmhDev P;
  memset(&P, 0, sizeof(P));
  P.C = C; P.p = p; P.nconst = nc; P.theta_stride = ts; P.seed = cfg->seed; P.chain_id_base = cfg->chain_id_base;
  for (int j = 0; j < p; j++) {
    P.prior_kind[j] = cfg->prior_kind[j]; P.prior_a[j] = cfg->prior_a[j]; P.prior_b[j] = cfg->prior_b[j];
    P.transform[j] = cfg->transform ? cfg->transform[j] : BSSM_TR_IDENTITY;
    P.pilot_sd[j] = cfg->pilot_proposal_sd ? cfg->pilot_proposal_sd[j] : 0.1;
  }
  for (int j = 0; j < nc; j++) P.consts[j] = cfg->consts[j];
  P.cur = B.cur; P.prop = B.prop; P.cur_ll = B.cur_ll; P.lp_prop = B.lp_prop; P.theta_full = B.theta_full;
  P.valid = B.valid; P.alive = B.alive; P.status = B.status; P.n_accept = B.n_accept; P.moved = d_moved;
  P.stream = B.ids; P.run_id = B.ids + C;

  // ---- filter batch for the chains ----
  FilterDev f;
  memset(&f, 0, sizeof(f));
  f.C = C; f.T = T; f.dy = cfg->dy; f.d = d; f.theta = B.theta_full; f.theta_stride = ts; f.y = B.y;
  f.obs_times = cfg->obs_times ? B.obs : nullptr; f.stream = P.stream; f.run_id = P.run_id; f.seed = cfg->seed;
  f.algorithm = cfg->algorithm; f.threshold = -1.0;
  FilterLaunch L;
  L.model = cfg->model; L.precision = cfg->precision; L.exact = (cfg->precision == BSSM_F64); L.hist = 0; L.T = T;
  L.engine = cfg->engine;
  const bool need_aux = cfg->algorithm == BSSM_APF;
  const int gb = (C + 127) / 128;
  double* cdf = nullptr;
  float pilot_ms = 0.f, main_ms = 0.f;
  std::vector<int> h_target(C, cfg->fixed_num_particles);

  BSSM_CK(cudaEventRecord(ctx->ev0, st));
  if (!cfg->skip_pilot) {
    // ---- pilot chain (R/pmmh_tuning.R:111-317) ----
    f.N = cfg->pilot_n; f.n_per = nullptr;
    f.ralg = cfg->algorithm == BSSM_RMPF ? BSSM_SISR : cfg->pilot_resample_algorithm;
    L.resample_fn = cfg->pilot_resample_fn;
    BSSM_TRY(filter_setup(ctx, f, L, need_aux, false, &cdf));
    P.f_loglike = f.loglike; P.f_status = f.status;
    k_pm_start<<<gb, 128, 0, st>>>(P, d_init, PH_PILOT, 1, 1);
    BSSM_LAUNCH(ctx, "k_pm_start");
    BSSM_TRY(run_chain_filters(ctx, f, L, cdf, P.valid));
    k_pm_first<<<gb, 128, 0, st>>>(P, B.pilot_chain, B.pilot_ll, pm);
    BSSM_LAUNCH(ctx, "k_pm_first");
    for (int it = 1; it < pm; it++) {
      k_pm_propose<<<gb, 128, 0, st>>>(P, PH_PILOT, it, nullptr);
      BSSM_LAUNCH(ctx, "k_pm_propose");
      BSSM_TRY(run_chain_filters(ctx, f, L, cdf, P.valid));
      k_pm_accept<<<gb, 128, 0, st>>>(P, PH_PILOT, it, B.pilot_chain, B.pilot_ll, pm);
      BSSM_LAUNCH(ctx, "k_pm_accept");
    }
    k_pm_pilot_stats<<<gb, 128, 0, st>>>(P, B.pilot_chain, pm, B.mean, B.cov);
    BSSM_LAUNCH(ctx, "k_pm_pilot_stats");
    // ---- .pilot_run: reps replicate filters per chain at the pilot mean.  They run with tune_control's pilot_resample_algorithm /
    //      pilot_resample_fn like the pilot chain: pmmh() hands both to .run_pilot_chain (R/pmmh.R:366-367), which has no such
    //      formals, so they travel in its `...` into do.call(.pilot_run, ...) (R/pmmh_tuning.R:292-305) and on to pf_wrapper
    //      (R/pmmh_tuning.R:34-50).  Only the MAIN chain runs on the wrapper's defaults (quirk A10) ----
    {
      FilterDev fr = f;
      fr.C = C * reps; fr.theta = B.theta_rep; fr.stream = B.ids_rep; fr.run_id = B.ids_rep + (size_t)C * reps;
      fr.ralg = cfg->algorithm == BSSM_RMPF ? BSSM_SISR : cfg->pilot_resample_algorithm;
      FilterLaunch Lr = L;
      Lr.resample_fn = cfg->pilot_resample_fn;
      BSSM_TRY(filter_setup(ctx, fr, Lr, need_aux, false, &cdf));
      k_pm_reps_setup<<<(C * reps + 127) / 128, 128, 0, st>>>(P, B.mean, reps, B.theta_rep, B.ids_rep, B.ids_rep + (size_t)C * reps, B.active_rep);
      BSSM_LAUNCH(ctx, "k_pm_reps_setup");
      BSSM_TRY(run_chain_filters(ctx, fr, Lr, cdf, B.active_rep));
      double* d_rep_ll;  // keep the replicate log-likelihoods in their own buffer for the result
      BSSM_TRY(scratch(ctx, slot++, (size_t)C * reps, &d_rep_ll));
      k_pm_tune<<<gb, 128, 0, st>>>(P, fr.loglike, fr.status, reps, cfg->pilot_n, cfg->fixed_num_particles, B.mean, B.cov,
                                    d_rep_ll, B.target_n, B.chol);
      BSSM_LAUNCH(ctx, "k_pm_tune");
      if (res->pilot_loglikes) BSSM_CK(cudaMemcpyAsync(res->pilot_loglikes, d_rep_ll, sizeof(double) * C * reps, cudaMemcpyDeviceToHost, st));
    }
    BSSM_CK(cudaMemcpyAsync(h_target.data(), B.target_n, sizeof(int) * C, cudaMemcpyDeviceToHost, st));
    BSSM_CK(cudaEventRecord(ctx->ev1, st));
    BSSM_CK(cudaStreamSynchronize(st));  // once per phase: target_n sizes the main-phase buffers
    BSSM_CK(cudaEventElapsedTime(&pilot_ms, ctx->ev0, ctx->ev1));
  } else {
    BSSM_CK(cudaMemcpyAsync(B.chol, cfg->proposal_chol_in, sizeof(double) * C * p * p, cudaMemcpyHostToDevice, st));
    BSSM_CK(cudaMemcpyAsync(B.mean, d_init, sizeof(double) * C * p, cudaMemcpyDeviceToDevice, st));
    BSSM_CK(cudaMemcpyAsync(B.target_n, h_target.data(), sizeof(int) * C, cudaMemcpyHostToDevice, st));
    BSSM_CK(cudaStreamSynchronize(st));
  }

  // ---- main chain (R/pmmh.R:395-500); filter defaults SISAR + stratified (quirk A10) ----
  int nmax = 1;
  for (int c = 0; c < C; c++) nmax = h_target[c] > nmax ? h_target[c] : nmax;
  bool ragged = false;
  for (int c = 0; c < C; c++) ragged = ragged || h_target[c] != nmax;
  f.C = C; f.N = nmax; f.n_per = ragged ? B.target_n : nullptr; f.theta = B.theta_full; f.stream = P.stream; f.run_id = P.run_id;
  f.ralg = cfg->algorithm == BSSM_RMPF ? BSSM_SISR : BSSM_SISAR;
  L.resample_fn = BSSM_STRATIFIED;
  BSSM_TRY(filter_setup(ctx, f, L, need_aux, false, &cdf));
  P.f_loglike = f.loglike; P.f_status = f.status; P.f_nres = f.n_resampled; P.nres_acc = d_nres;
  BSSM_CK(cudaEventRecord(ctx->ev0, st));
  k_pm_start<<<gb, 128, 0, st>>>(P, B.mean, PH_MAIN, cfg->skip_pilot ? 1 : 0, cfg->skip_pilot ? 1 : 0);
  BSSM_LAUNCH(ctx, "k_pm_start");
  BSSM_TRY(run_chain_filters(ctx, f, L, cdf, P.valid));
  k_pm_first<<<gb, 128, 0, st>>>(P, B.chain, B.ll_chain, m);
  BSSM_LAUNCH(ctx, "k_pm_first");
  const dim3 grid_se((se_len + 255) / 256 > 64 ? 64 : (se_len + 255) / 256, C);
  if (latent) { k_pm_latent<<<grid_se, 256, 0, st>>>(P, f.state_est, se_len, d_cur_se, d_se_chain, 0, m); BSSM_LAUNCH(ctx, "k_pm_latent"); }
  for (int it = 1; it < m; it++) {
    k_pm_propose<<<gb, 128, 0, st>>>(P, PH_MAIN, it, B.chol);
    BSSM_LAUNCH(ctx, "k_pm_propose");
    BSSM_TRY(run_chain_filters(ctx, f, L, cdf, P.valid));
    k_pm_accept<<<gb, 128, 0, st>>>(P, PH_MAIN, it, B.chain, B.ll_chain, m);
    BSSM_LAUNCH(ctx, "k_pm_accept");
    if (latent) { k_pm_latent<<<grid_se, 256, 0, st>>>(P, f.state_est, se_len, d_cur_se, d_se_chain, it, m); BSSM_LAUNCH(ctx, "k_pm_latent"); }
  }
  BSSM_CK(cudaEventRecord(ctx->ev1, st));

  // ---- results ----
#define DL(dst, src, count, type) if (dst) BSSM_CK(cudaMemcpyAsync(dst, src, (count) * sizeof(type), cudaMemcpyDeviceToHost, st))
  if (!cfg->skip_pilot) {
    DL(res->pilot_theta_chain, B.pilot_chain, (size_t)C * pm * p, double);
    DL(res->pilot_loglike_chain, B.pilot_ll, (size_t)C * pm, double);
    DL(res->pilot_theta_cov, B.cov, (size_t)C * p * p, double);
  }
  DL(res->pilot_theta_mean, B.mean, (size_t)C * p, double);
  DL(res->target_n, B.target_n, (size_t)C, int);
  DL(res->proposal_chol, B.chol, (size_t)C * p * p, double);
  DL(res->theta_chain, B.chain, (size_t)C * m * p, double);
  DL(res->loglike_chain, B.ll_chain, (size_t)C * m, double);
  DL(res->n_accept, B.n_accept, (size_t)C, int);
  if (latent) DL(res->latent_state_chain, d_se_chain, (size_t)C * m * se_len, double);
  DL(res->status, B.status, (size_t)C, int);
  unsigned long long h_nres[2] = {0, 0};
  BSSM_CK(cudaMemcpyAsync(h_nres, d_nres, sizeof(h_nres), cudaMemcpyDeviceToHost, st));
#undef DL
  BSSM_CK(cudaStreamSynchronize(st));
  BSSM_CK(cudaEventElapsedTime(&main_ms, ctx->ev0, ctx->ev1));
  res->pilot_ms = pilot_ms;
  res->main_ms = main_ms;
  res->main_resampled_fraction = (h_nres[1] && T) ? (double)h_nres[0] / ((double)h_nres[1] * T) : 0.0;
  return BSSM_OK;
}

double bssm_transform(double th, int tr) { return tr == BSSM_TR_LOG ? log(th) : (tr == BSSM_TR_LOGIT ? log(th / (1.0 - th)) : th); }
double bssm_back_transform(double z, int tr) { return tr == BSSM_TR_LOG ? exp(z) : (tr == BSSM_TR_LOGIT ? 1.0 / (1.0 + exp(-z)) : z); }
double bssm_log_jacobian(const double* th, const int* tr, int p) {
  double s = 0;
  for (int j = 0; j < p; j++) {
    if (tr[j] == BSSM_TR_LOG) s += log(th[j]);
    else if (tr[j] == BSSM_TR_LOGIT) s += log(1.0 / (th[j] * (1.0 - th[j])));
  }
  return s;
}
double bssm_log_prior(int kind, double a, double b, double x) {
  const double LSP = 0.918938533204672741780329736406;
  switch (kind) {
    case BSSM_PRIOR_FLAT: return 0.0;
    case BSSM_PRIOR_NORMAL: { double z = (x - a) / b; return -(LSP + 0.5 * z * z + log(b)); }
    case BSSM_PRIOR_EXP: return x < 0 ? -INFINITY : log(a) - a * x;
    case BSSM_PRIOR_UNIF: return (a <= x && x <= b) ? -log(b - a) : -INFINITY;
    case BSSM_PRIOR_HALFNORMAL: { if (x < 0) return -INFINITY; double z = x / a; return log(2.0) - (LSP + 0.5 * z * z + log(a)); }
  }
  return NAN;
}

}  // extern "C"
