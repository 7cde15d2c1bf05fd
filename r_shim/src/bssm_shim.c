/*
 * bssm_shim.c -- thin .Call shim binding libbayesssm_b200.so (include/bayesssm_b200.h) into R.
 *
 * Replaces src/RcppExports.cpp of bayesSSM (generated Rcpp glue, :15-60) by hand: same three
 * registered symbols with the same arity, plus the filter / PMMH entry points.  Logic-free on
 * purpose: argument unpacking, one C-ABI call, result packing.  R and its headers are not
 * installed in the build image, so this file is compile-checked only where R exists
 * (R CMD SHLIB bssm_shim.c -L... -lbayesssm_b200); see INTEGRATION.md.
 *
 * Rules kept: no SEXP is stored; every R allocation is PROTECTed until return; the engine
 * never throws across the ABI, so Rf_error() is only called after the C call has returned and
 * nothing native is left to release; uniforms are drawn with unif_rand() between
 * GetRNGstate()/PutRNGstate() so set.seed() keeps its meaning (src/RcppExports.cpp:18,30,42).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "bayesssm_b200.h"

static bssm_ctx *g_ctx = NULL;

static bssm_ctx *ctx_get(void) {
  if (!g_ctx) {
    int st = bssm_create(0, &g_ctx);
    if (st != BSSM_OK) Rf_error("bayesSSM (B200 engine): %s", bssm_last_error());
  }
  return g_ctx;
}

static SEXP resample_common(SEXP n_, SEXP weights_, int kind) {
  int n = Rf_asInteger(n_);
  if (TYPEOF(weights_) != REALSXP) Rf_error("weights must be numeric");
  if (XLENGTH(weights_) != n) Rf_error("Length of weights must match n");
  bssm_ctx *ctx = ctx_get();
  int nu = kind == BSSM_SYSTEMATIC ? 1 : n;
  double *u = (double *)R_alloc((size_t)nu, sizeof(double));
  GetRNGstate();
  for (int i = 0; i < nu; i++) u[i] = unif_rand(); /* Rcpp::runif(n) / R::runif(0,1), src/resampling.cpp:28,55 */
  PutRNGstate();
  SEXP out = PROTECT(Rf_allocVector(INTSXP, n));
  int st;
  if (kind == BSSM_STRATIFIED) st = bssm_resample_stratified(ctx, n, REAL(weights_), u, INTEGER(out));
  else if (kind == BSSM_SYSTEMATIC) st = bssm_resample_systematic(ctx, n, REAL(weights_), u[0], INTEGER(out));
  else st = bssm_resample_multinomial(ctx, n, REAL(weights_), u, INTEGER(out));
  UNPROTECT(1);
  if (st != BSSM_OK) Rf_error("%s", bssm_last_error()); /* "Weights must be non-negative" / "Sum of weights must be greater than 0" */
  return out;
}
/* the three routines R/RcppExports.R:4-14 calls */
SEXP _bayesSSM_resample_multinomial_cpp(SEXP n, SEXP w) { return resample_common(n, w, BSSM_MULTINOMIAL); }
SEXP _bayesSSM_resample_stratified_cpp(SEXP n, SEXP w) { return resample_common(n, w, BSSM_STRATIFIED); }
SEXP _bayesSSM_resample_systematic_cpp(SEXP n, SEXP w) { return resample_common(n, w, BSSM_SYSTEMATIC); }

static SEXP list_get(SEXP list, const char *name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  for (R_xlen_t i = 0; i < XLENGTH(list); i++)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  return R_NilValue;
}
static int opt_int(SEXP cfg, const char *name, int dflt) {
  SEXP v = list_get(cfg, name);
  return v == R_NilValue ? dflt : Rf_asInteger(v);
}
static double opt_real(SEXP cfg, const char *name, double dflt) {
  SEXP v = list_get(cfg, name);
  return v == R_NilValue ? dflt : Rf_asReal(v);
}

/* .Call("_bayesSSM_b200_filter", cfg (named list of scalars), y (T x dy matrix, column-major), theta)
 * -> list(state_est, ess, loglike, loglike_history, early_exit, n_resampled[, particles_history, weights_history]) */
SEXP _bayesSSM_b200_filter(SEXP cfg_, SEXP y_, SEXP theta_) {
  bssm_ctx *ctx = ctx_get();
  bssm_filter_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.model = opt_int(cfg_, "model", 0);
  cfg.algorithm = opt_int(cfg_, "algorithm", BSSM_BPF);
  cfg.resample_algorithm = opt_int(cfg_, "resample_algorithm", BSSM_SISAR);
  cfg.resample_fn = opt_int(cfg_, "resample_fn", BSSM_STRATIFIED);
  cfg.carry_weights = opt_int(cfg_, "carry_weights", 0);   /* extension: standard SMC weights (not the reference's rule) */
  cfg.threshold = opt_real(cfg_, "threshold", -1.0);
  cfg.num_particles = opt_int(cfg_, "num_particles", 0);
  cfg.num_obs = Rf_isMatrix(y_) ? Rf_nrows(y_) : (int)XLENGTH(y_);
  cfg.dy = Rf_isMatrix(y_) ? Rf_ncols(y_) : 1;
  cfg.num_filters = 1;
  cfg.precision = opt_int(cfg_, "precision", BSSM_F64);
  cfg.seed = (uint64_t)opt_real(cfg_, "seed", 1.0);
  cfg.return_particles = opt_int(cfg_, "return_particles", 1);
  cfg.exact_resampling = -1;
  cfg.engine = opt_int(cfg_, "engine", BSSM_ENGINE_AUTO);
  SEXP ot = list_get(cfg_, "obs_times");
  if (ot != R_NilValue) cfg.obs_times = INTEGER(ot);
  int d = 1, nth = 0, nc = 0;
  if (bssm_model_dims(ctx, cfg.model, &d, &nth, &nc) != BSSM_OK) Rf_error("%s", bssm_last_error());
  if (XLENGTH(theta_) != nth + nc) Rf_error("model expects %d parameters and %d constants", nth, nc);
  const int T = cfg.num_obs, N = cfg.num_particles;
  /* R matrices are column-major; the engine wants y[T][dy] row-major */
  double *y = (double *)R_alloc((size_t)T * cfg.dy, sizeof(double));
  for (int t = 0; t < T; t++) for (int k = 0; k < cfg.dy; k++) y[(size_t)t * cfg.dy + k] = REAL(y_)[(size_t)k * T + t];
  SEXP state_est = PROTECT(d == 1 ? Rf_allocVector(REALSXP, T + 1) : Rf_allocMatrix(REALSXP, T + 1, d));
  SEXP ess = PROTECT(Rf_allocVector(REALSXP, T + 1));
  SEXP llh = PROTECT(Rf_allocVector(REALSXP, T));
  double *se_rm = (double *)R_alloc((size_t)(T + 1) * d, sizeof(double));
  double loglike = 0.0;
  int32_t status = 0, early = 0, nres = 0;
  bssm_filter_result res;
  memset(&res, 0, sizeof(res));
  res.loglike = &loglike; res.loglike_history = REAL(llh); res.ess = REAL(ess); res.state_est = se_rm;
  res.status = &status; res.early_exit = &early; res.n_resampled = &nres;
  SEXP ph = R_NilValue, wh = R_NilValue;
  int nprot = 3;
  double *ph_rm = NULL;
  if (cfg.return_particles) {
    ph = PROTECT(Rf_allocMatrix(REALSXP, T + 1, N * d)); /* rbind(as.numeric(particles)): R/particle_filter_core.R:256-262 */
    wh = PROTECT(Rf_allocMatrix(REALSXP, T + 1, N));
    nprot += 2;
    ph_rm = (double *)R_alloc((size_t)(T + 1) * d * N, sizeof(double));
    res.particles_history = ph_rm;
    res.weights_history = (double *)R_alloc((size_t)(T + 1) * N, sizeof(double));
  }
  int st = bssm_filter_run(ctx, &cfg, y, REAL(theta_), &res);
  if (st != BSSM_OK) { UNPROTECT(nprot); Rf_error("%s", bssm_last_error()); }
  if (status == BSSM_ERR_NAN_WEIGHT) { UNPROTECT(nprot); Rf_error("missing value where TRUE/FALSE needed"); }
  for (int t = 0; t <= T; t++) for (int k = 0; k < d; k++) REAL(state_est)[(size_t)k * (T + 1) + t] = se_rm[(size_t)t * d + k];
  if (cfg.return_particles) {
    for (int t = 0; t <= T; t++) {
      for (size_t j = 0; j < (size_t)N * d; j++) REAL(ph)[j * (T + 1) + t] = ph_rm[(size_t)t * d * N + j];
      for (int j = 0; j < N; j++) REAL(wh)[(size_t)j * (T + 1) + t] = res.weights_history[(size_t)t * N + j];
    }
  }
  const char *names[] = {"state_est", "ess", "loglike", "loglike_history", "early_exit", "n_resampled",
                         "particles_history", "weights_history", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  nprot++;
  SET_VECTOR_ELT(out, 0, state_est); SET_VECTOR_ELT(out, 1, ess);
  SET_VECTOR_ELT(out, 2, Rf_ScalarReal(loglike)); SET_VECTOR_ELT(out, 3, llh);
  SET_VECTOR_ELT(out, 4, Rf_ScalarLogical(early)); SET_VECTOR_ELT(out, 5, Rf_ScalarInteger(nres));
  SET_VECTOR_ELT(out, 6, ph); SET_VECTOR_ELT(out, 7, wh);
  UNPROTECT(nprot);
  return out;
}

/* .Call("_bayesSSM_b200_pmmh", cfg (named list), y, init_theta (num_chains x p matrix))
 * -> list(theta_chain [chains][m][p] as an array, loglike_chain, target_n, n_accept, status, pilot_theta_mean, ...) */
SEXP _bayesSSM_b200_pmmh(SEXP cfg_, SEXP y_, SEXP init_) {
  bssm_ctx *ctx = ctx_get();
  bssm_pmmh_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  const int C = Rf_nrows(init_), p = Rf_ncols(init_);
  cfg.model = opt_int(cfg_, "model", 0); cfg.algorithm = opt_int(cfg_, "algorithm", BSSM_BPF); cfg.p = p;
  cfg.prior_kind = INTEGER(list_get(cfg_, "prior_kind")); cfg.prior_a = REAL(list_get(cfg_, "prior_a"));
  cfg.prior_b = REAL(list_get(cfg_, "prior_b")); cfg.transform = INTEGER(list_get(cfg_, "transform"));
  cfg.pilot_proposal_sd = REAL(list_get(cfg_, "pilot_proposal_sd"));
  cfg.pilot_n = opt_int(cfg_, "pilot_n", 100); cfg.pilot_m = opt_int(cfg_, "pilot_m", 2000);
  cfg.pilot_reps = opt_int(cfg_, "pilot_reps", 100);
  cfg.pilot_resample_algorithm = opt_int(cfg_, "pilot_resample_algorithm", BSSM_SISAR);
  cfg.pilot_resample_fn = opt_int(cfg_, "pilot_resample_fn", BSSM_STRATIFIED);
  cfg.m = opt_int(cfg_, "m", 1000); cfg.num_chains = C; cfg.chain_id_base = 0;
  cfg.fixed_num_particles = opt_int(cfg_, "num_particles", 0);
  cfg.num_obs = Rf_isMatrix(y_) ? Rf_nrows(y_) : (int)XLENGTH(y_);
  cfg.dy = Rf_isMatrix(y_) ? Rf_ncols(y_) : 1;
  SEXP ot = list_get(cfg_, "obs_times");
  if (ot != R_NilValue) cfg.obs_times = INTEGER(ot);
  SEXP cs = list_get(cfg_, "consts");
  if (cs != R_NilValue) { cfg.consts = REAL(cs); cfg.nconst = (int)XLENGTH(cs); }
  cfg.precision = opt_int(cfg_, "precision", BSSM_F64);
  cfg.seed = (uint64_t)opt_real(cfg_, "seed", 1.0);
  cfg.engine = opt_int(cfg_, "engine", BSSM_ENGINE_AUTO);
  const int T = cfg.num_obs, m = cfg.m;
  double *y = (double *)R_alloc((size_t)T * cfg.dy, sizeof(double));
  for (int t = 0; t < T; t++) for (int k = 0; k < cfg.dy; k++) y[(size_t)t * cfg.dy + k] = REAL(y_)[(size_t)k * T + t];
  double *init = (double *)R_alloc((size_t)C * p, sizeof(double));
  for (int c = 0; c < C; c++) for (int j = 0; j < p; j++) init[(size_t)c * p + j] = REAL(init_)[(size_t)j * C + c];
  bssm_pmmh_result res;
  memset(&res, 0, sizeof(res));
  res.theta_chain = (double *)R_alloc((size_t)C * m * p, sizeof(double));
  SEXP ll = PROTECT(Rf_allocMatrix(REALSXP, m, C));
  SEXP tn = PROTECT(Rf_allocVector(INTSXP, C)), na = PROTECT(Rf_allocVector(INTSXP, C)), stv = PROTECT(Rf_allocVector(INTSXP, C));
  SEXP pmean = PROTECT(Rf_allocMatrix(REALSXP, C, p));
  double *ll_rm = (double *)R_alloc((size_t)C * m, sizeof(double)), *pm_rm = (double *)R_alloc((size_t)C * p, sizeof(double));
  res.loglike_chain = ll_rm; res.target_n = INTEGER(tn); res.n_accept = INTEGER(na); res.status = INTEGER(stv);
  res.pilot_theta_mean = pm_rm;
  int dstate = 1, nth_ = 0, nc_ = 0;
  if (bssm_model_dims(ctx, cfg.model, &dstate, &nth_, &nc_) != BSSM_OK) { UNPROTECT(5); Rf_error("%s", bssm_last_error()); }
  cfg.return_latent_state_est = opt_int(cfg_, "return_latent_state_est", 0);
  const size_t se_len = (size_t)(T + 1) * dstate;
  if (cfg.return_latent_state_est) res.latent_state_chain = (double *)R_alloc((size_t)C * m * se_len, sizeof(double));
  int st = bssm_pmmh_run(ctx, &cfg, y, init, &res);
  if (st != BSSM_OK) { UNPROTECT(5); Rf_error("%s", bssm_last_error()); }
  SEXP dims = PROTECT(Rf_allocVector(INTSXP, 3));
  INTEGER(dims)[0] = m; INTEGER(dims)[1] = p; INTEGER(dims)[2] = C;
  SEXP th = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)m * p * C)); /* array [m, p, chain] */
  for (int c = 0; c < C; c++) for (int i = 0; i < m; i++) for (int j = 0; j < p; j++)
    REAL(th)[((size_t)c * p + j) * m + i] = res.theta_chain[((size_t)c * m + i) * p + j];
  Rf_setAttrib(th, R_DimSymbol, dims);
  for (int c = 0; c < C; c++) {
    for (int i = 0; i < m; i++) REAL(ll)[(size_t)c * m + i] = ll_rm[(size_t)c * m + i];
    for (int j = 0; j < p; j++) REAL(pmean)[(size_t)j * C + c] = pm_rm[(size_t)c * p + j];
  }
  /* latent_state_chain: array [T+1, d, m, chain] (R/pmmh.R:420,494-499), or NULL */
  SEXP lat = R_NilValue;
  int nprot = 8;
  if (cfg.return_latent_state_est) {
    lat = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)((size_t)C * m * se_len)));
    nprot++;
    for (int c = 0; c < C; c++) for (int i = 0; i < m; i++) for (int t = 0; t <= T; t++) for (int k = 0; k < dstate; k++)
      REAL(lat)[(((size_t)c * m + i) * dstate + k) * (T + 1) + t] = res.latent_state_chain[(((size_t)c * m + i) * (T + 1) + t) * dstate + k];
    SEXP ld = PROTECT(Rf_allocVector(INTSXP, 4));
    nprot++;
    INTEGER(ld)[0] = T + 1; INTEGER(ld)[1] = dstate; INTEGER(ld)[2] = m; INTEGER(ld)[3] = C;
    Rf_setAttrib(lat, R_DimSymbol, ld);
  }
  const char *names[] = {"theta_chain", "loglike_chain", "target_n", "n_accept", "status", "pilot_theta_mean", "latent_state_chain", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(out, 0, th); SET_VECTOR_ELT(out, 1, ll); SET_VECTOR_ELT(out, 2, tn);
  SET_VECTOR_ELT(out, 3, na); SET_VECTOR_ELT(out, 4, stv); SET_VECTOR_ELT(out, 5, pmean); SET_VECTOR_ELT(out, 6, lat);
  UNPROTECT(nprot);
  return out;
}

/* ---- particle-sharded single filter: one R process per GPU (e.g. under mpirun / callr workers) ----
 * .Call("_bayesSSM_b200_shard_unique_id")                  rank 0: raw(128), to be sent to the other ranks
 * .Call("_bayesSSM_b200_shard_init", rank, world, id)      every rank (device = rank of the process on the node)
 * .Call("_bayesSSM_b200_shard_filter", cfg, y, theta)      collective; cfg$num_particles is the GLOBAL count */
SEXP _bayesSSM_b200_shard_unique_id(void) {
  SEXP id = PROTECT(Rf_allocVector(RAWSXP, 128));
  int st = bssm_shard_unique_id(NULL, RAW(id));
  UNPROTECT(1);
  if (st != BSSM_OK) Rf_error("%s", bssm_last_error());
  return id;
}
SEXP _bayesSSM_b200_shard_init(SEXP rank_, SEXP world_, SEXP id_) {
  const int rank = Rf_asInteger(rank_), world = Rf_asInteger(world_);
  if (!g_ctx) { if (bssm_create(rank, &g_ctx) != BSSM_OK) Rf_error("bayesSSM (B200 engine): %s", bssm_last_error()); }
  if (world > 1 && (TYPEOF(id_) != RAWSXP || XLENGTH(id_) != 128)) Rf_error("id must be the raw(128) of shard_unique_id");
  if (bssm_shard_init(g_ctx, NULL, rank, world, world > 1 ? RAW(id_) : NULL) != BSSM_OK) Rf_error("%s", bssm_last_error());
  return R_NilValue;
}
/* optional: the per-observation exchange through peer memory instead of ncclAllGather (include/bayesssm_b200.h, bssm_shard_peer_*).
 * .Call("_bayesSSM_b200_shard_peer_export")           every rank, after shard_init: raw(64), to be gathered in rank order
 * .Call("_bayesSSM_b200_shard_peer_attach", handles)  every rank: raw(64 * world), the handles of ranks 0 .. world - 1 */
SEXP _bayesSSM_b200_shard_peer_export(void) {
  SEXP h = PROTECT(Rf_allocVector(RAWSXP, 64));
  int st = bssm_shard_peer_export(ctx_get(), RAW(h));
  UNPROTECT(1);
  if (st != BSSM_OK) Rf_error("%s", bssm_last_error());
  return h;
}
SEXP _bayesSSM_b200_shard_peer_attach(SEXP handles_) {
  if (TYPEOF(handles_) != RAWSXP || XLENGTH(handles_) % 64 != 0) Rf_error("handles must be the concatenated raw(64) of every rank, in rank order");
  if (bssm_shard_peer_attach(ctx_get(), RAW(handles_)) != BSSM_OK) Rf_error("%s", bssm_last_error());
  return R_NilValue;
}
SEXP _bayesSSM_b200_shard_filter(SEXP cfg_, SEXP y_, SEXP theta_) {
  bssm_ctx *ctx = ctx_get();
  bssm_filter_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.model = opt_int(cfg_, "model", 0);
  cfg.algorithm = BSSM_BPF;
  cfg.resample_algorithm = opt_int(cfg_, "resample_algorithm", BSSM_SISAR);
  cfg.resample_fn = opt_int(cfg_, "resample_fn", BSSM_STRATIFIED);
  cfg.threshold = opt_real(cfg_, "threshold", -1.0);
  cfg.num_particles = opt_int(cfg_, "num_particles", 0);
  cfg.num_obs = Rf_isMatrix(y_) ? Rf_nrows(y_) : (int)XLENGTH(y_);
  cfg.dy = Rf_isMatrix(y_) ? Rf_ncols(y_) : 1;
  cfg.num_filters = 1;
  cfg.precision = opt_int(cfg_, "precision", BSSM_F32);
  cfg.seed = (uint64_t)opt_real(cfg_, "seed", 1.0);
  cfg.engine = BSSM_ENGINE_STREAM;
  const int T = cfg.num_obs;
  double *y = (double *)R_alloc((size_t)T * cfg.dy, sizeof(double));
  for (int t = 0; t < T; t++) for (int k = 0; k < cfg.dy; k++) y[(size_t)t * cfg.dy + k] = REAL(y_)[(size_t)k * T + t];
  SEXP state_est = PROTECT(Rf_allocVector(REALSXP, T + 1)), ess = PROTECT(Rf_allocVector(REALSXP, T + 1));
  SEXP llh = PROTECT(Rf_allocVector(REALSXP, T));
  double loglike = 0.0;
  int32_t status = 0, early = 0, nres = 0;
  int nloc = 0;
  bssm_filter_result res;
  memset(&res, 0, sizeof(res));
  res.loglike = &loglike; res.loglike_history = REAL(llh); res.ess = REAL(ess); res.state_est = REAL(state_est);
  res.status = &status; res.early_exit = &early; res.n_resampled = &nres;
  int st = bssm_filter_run_sharded(ctx, &cfg, y, REAL(theta_), opt_real(cfg_, "capacity_factor", 1.5), &res, &nloc);
  if (st != BSSM_OK) { UNPROTECT(3); Rf_error("%s", bssm_last_error()); }
  if (status == BSSM_ERR_CAPACITY) { UNPROTECT(3); Rf_error("a rank's share of the offspring outgrew its storage; raise capacity_factor"); }
  if (status == BSSM_ERR_NAN_WEIGHT) { UNPROTECT(3); Rf_error("missing value where TRUE/FALSE needed"); }
  const char *names[] = {"state_est", "ess", "loglike", "loglike_history", "early_exit", "n_resampled", "n_local", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(out, 0, state_est); SET_VECTOR_ELT(out, 1, ess);
  SET_VECTOR_ELT(out, 2, Rf_ScalarReal(loglike)); SET_VECTOR_ELT(out, 3, llh);
  SET_VECTOR_ELT(out, 4, Rf_ScalarLogical(early)); SET_VECTOR_ELT(out, 5, Rf_ScalarInteger(nres));
  SET_VECTOR_ELT(out, 6, Rf_ScalarInteger(nloc));
  UNPROTECT(4);
  return out;
}

/* .Call("_bayesSSM_b200_model_compile", source) -> list(model, d, ntheta, nconst): a CUDA snippet defining
 * `struct UserModel` (contract: bayesssm_b200/csrc/bssm_models.cuh) compiled by NVRTC for sm_100a; stands where the
 * reference takes R closures (R/particle_filter-doc.R:11-18). */
SEXP _bayesSSM_b200_model_compile(SEXP src_) {
  if (TYPEOF(src_) != STRSXP || XLENGTH(src_) != 1) Rf_error("source must be a character(1) CUDA snippet");
  bssm_ctx *ctx = ctx_get();
  int id = 0, d = 0, nth = 0, nc = 0;
  if (bssm_model_compile(ctx, CHAR(STRING_ELT(src_, 0)), &id) != BSSM_OK) Rf_error("%s", bssm_last_error());
  if (bssm_model_dims(ctx, id, &d, &nth, &nc) != BSSM_OK) Rf_error("%s", bssm_last_error());
  const char *names[] = {"model", "d", "ntheta", "nconst", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(out, 0, Rf_ScalarInteger(id)); SET_VECTOR_ELT(out, 1, Rf_ScalarInteger(d));
  SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(nth)); SET_VECTOR_ELT(out, 3, Rf_ScalarInteger(nc));
  UNPROTECT(1);
  return out;
}

SEXP _bayesSSM_b200_device_info(void) {
  char name[256];
  int sm = 0, maj = 0, mnr = 0;
  size_t mem = 0;
  if (bssm_device_info(ctx_get(), name, &sm, &maj, &mnr, &mem) != BSSM_OK) Rf_error("%s", bssm_last_error());
  return Rf_mkString(name);
}

/* ess() / rhat() of one parameter (R/ESS.R:30-104, R/rhat.R:27-67): `mat` is the reference's m x k REAL matrix
 * (iterations x chains); column-major storage is exactly draws[k][m][1] of bssm_mcmc_diagnostics.  Returns
 * list(ess, rhat, flags); the R wrapper turns flags into NA + warning("One or more chains have zero variance."). */
SEXP _bayesSSM_b200_mcmc_diagnostics(SEXP mat, SEXP want_ess_) {
  if (!Rf_isMatrix(mat)) Rf_error("Input must be a matrix or a data frame with a 'chain' column.");
  const int m = Rf_nrows(mat), k = Rf_ncols(mat), want_ess = Rf_asInteger(want_ess_);
  double ess = NA_REAL, rhat = NA_REAL;
  int32_t flags = 0;
  if (bssm_mcmc_diagnostics(ctx_get(), REAL(mat), k, m, 1, 0, want_ess ? &ess : NULL, &rhat, &flags, NULL) != BSSM_OK)
    Rf_error("%s", bssm_last_error()); /* "Number of iterations / chains must be at least 2." */
  const char *names[] = {"ess", "rhat", "flags", ""};
  SEXP out = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(out, 0, Rf_ScalarReal((flags & 1) || !want_ess ? NA_REAL : ess));
  SET_VECTOR_ELT(out, 1, Rf_ScalarReal((flags & 2) ? NA_REAL : rhat));
  SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(flags));
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"_bayesSSM_resample_multinomial_cpp", (DL_FUNC)&_bayesSSM_resample_multinomial_cpp, 2},
    {"_bayesSSM_resample_stratified_cpp", (DL_FUNC)&_bayesSSM_resample_stratified_cpp, 2},
    {"_bayesSSM_resample_systematic_cpp", (DL_FUNC)&_bayesSSM_resample_systematic_cpp, 2},
    {"_bayesSSM_b200_filter", (DL_FUNC)&_bayesSSM_b200_filter, 3},
    {"_bayesSSM_b200_pmmh", (DL_FUNC)&_bayesSSM_b200_pmmh, 3},
    {"_bayesSSM_b200_device_info", (DL_FUNC)&_bayesSSM_b200_device_info, 0},
    {"_bayesSSM_b200_model_compile", (DL_FUNC)&_bayesSSM_b200_model_compile, 1},
    {"_bayesSSM_b200_mcmc_diagnostics", (DL_FUNC)&_bayesSSM_b200_mcmc_diagnostics, 2},
    {"_bayesSSM_b200_shard_unique_id", (DL_FUNC)&_bayesSSM_b200_shard_unique_id, 0},
    {"_bayesSSM_b200_shard_init", (DL_FUNC)&_bayesSSM_b200_shard_init, 3},
    {"_bayesSSM_b200_shard_filter", (DL_FUNC)&_bayesSSM_b200_shard_filter, 3},
    {"_bayesSSM_b200_shard_peer_export", (DL_FUNC)&_bayesSSM_b200_shard_peer_export, 0},
    {"_bayesSSM_b200_shard_peer_attach", (DL_FUNC)&_bayesSSM_b200_shard_peer_attach, 1},
    {NULL, NULL, 0}};

void R_init_bayesSSM(DllInfo *dll) { /* replaces src/RcppExports.cpp:57-60 */
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
void R_unload_bayesSSM(DllInfo *dll) {
  (void)dll;
  if (g_ctx) { bssm_shard_finalize(g_ctx); bssm_destroy(g_ctx); g_ctx = NULL; }
}
