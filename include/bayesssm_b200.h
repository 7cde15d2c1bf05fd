/*
 * bayesssm_b200.h -- C ABI of the B200-native particle-filter / PMMH engine.
 *
 * Drop-in boundary for the hot path of the R package bayesSSM (0.7.1.9000).
 * Plain C types only: pointers, sizes, POD structs, int status codes.  No C++
 * exception crosses this boundary; on failure a function returns a non-zero
 * BSSM_ERR_* and bssm_last_error() returns a thread-local message.  Caller
 * owns every host buffer; device memory is owned by the opaque context.
 *
 * Each entry point names the reference interface it replaces (paths relative
 * to the bayesSSM source tree).  The R-side `.Call` shim that binds these is
 * in r_shim/ and described in INTEGRATION.md.
 *
 * There is no CPU fallback: every function needs a CUDA device (sm_100a).
 */
#ifndef BAYESSSM_B200_H
#define BAYESSSM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSSM_ABI_VERSION 1

/* ---- status codes ---- */
#define BSSM_OK 0
#define BSSM_ERR_NEGATIVE_WEIGHT 1 /* R: "Weights must be non-negative"          src/resampling.cpp:6,18,45 */
#define BSSM_ERR_ZERO_SUM 2        /* R: "Sum of weights must be greater than 0" src/resampling.cpp:8,22,49 */
#define BSSM_ERR_NAN_WEIGHT 3      /* R: "missing value where TRUE/FALSE needed" R/particle_filter_core.R:189 */
#define BSSM_ERR_BAD_ARG 4
#define BSSM_ERR_PRIOR_INIT 5      /* R: "Initial parameter values are invalid ..." R/pmmh_tuning.R:136-142 */
#define BSSM_ERR_CUDA 6
#define BSSM_ERR_NVRTC 7
#define BSSM_ERR_UNSUPPORTED 8
#define BSSM_ERR_NO_DEVICE 9
#define BSSM_ERR_CAPACITY 10       /* particle-sharded filter: one rank's share of the offspring outgrew its storage */
#define BSSM_ERR_NCCL 11

/* ---- enums (R match.arg strings -> ints) ---- */
enum { BSSM_BPF = 0, BSSM_APF = 1, BSSM_RMPF = 2 };            /* pf_wrapper identity */
enum { BSSM_SIS = 0, BSSM_SISR = 1, BSSM_SISAR = 2 };          /* resample_algorithm  */
enum { BSSM_STRATIFIED = 0, BSSM_SYSTEMATIC = 1, BSSM_MULTINOMIAL = 2 }; /* resample_fn */
enum { BSSM_F32 = 0, BSSM_F64 = 1 };                          /* state / weight precision (cdf is always f64) */
enum { BSSM_ENGINE_AUTO = 0, BSSM_ENGINE_GENERAL = 1, BSSM_ENGINE_PERSISTENT = 2, BSSM_ENGINE_STREAM = 3 };
enum {
  BSSM_MODEL_AR_SIN = 0,   /* README.md:137-146                       theta = (phi, sigma_x, sigma_y) */
  BSSM_MODEL_LG = 1,       /* tests/testthat/test-pmmh_tuning.R:163   theta = (phi, sigma_x, sigma_y) */
  BSSM_MODEL_RW_DRIFT = 2, /* tests/testthat/test-auxiliary_filter.R  theta = (mu, sigma)             */
  BSSM_MODEL_SIR_CB = 3,   /* chain-binomial SIR, Poisson obs         theta = (lambda, gamma), consts = (pop, I0) */
  BSSM_MODEL_SIR_GILLESPIE = 6, /* the same model with the exact (Gillespie) daily step of vignettes/articles/stochastic-sir-model.Rmd:152-176;
                               Philox noise only (a data-dependent number of uniforms per transition) */
  BSSM_MODEL_AR_COS = 4,   /* R/pmmh.R:157-159                         theta = (phi, sigma_x, sigma_y) */
  BSSM_MODEL_RW2D = 5,     /* tests/testthat/test-bootstrap_filter.R:211  theta = (phi)               */
  BSSM_MODEL_BUILTIN_COUNT = 6
  /* ids >= 1000 are NVRTC-compiled user models returned by bssm_model_compile() */
};
enum { BSSM_PRIOR_FLAT = 0, BSSM_PRIOR_NORMAL = 1, BSSM_PRIOR_EXP = 2, BSSM_PRIOR_UNIF = 3, BSSM_PRIOR_HALFNORMAL = 4 };
enum { BSSM_TR_IDENTITY = 0, BSSM_TR_LOG = 1, BSSM_TR_LOGIT = 2 };     /* R/utils.R:102-152 */

typedef struct bssm_ctx bssm_ctx;

/* ---- context ---- */
int bssm_abi_version(void);
const char *bssm_last_error(void);
int bssm_create(int device, bssm_ctx **out);
void bssm_destroy(bssm_ctx *ctx);
/* name buffer >= 256 bytes */
int bssm_device_info(bssm_ctx *ctx, char *name, int *sm_count, int *cc_major, int *cc_minor, size_t *global_mem);
/* number of engine kernel launches issued by this context so far (bench.py gpu_launches) */
int64_t bssm_launch_count(bssm_ctx *ctx);
int bssm_synchronize(bssm_ctx *ctx);
/* CUDA-event timing on the context's stream (used by bench.py) */
int bssm_timer_start(bssm_ctx *ctx);
int bssm_timer_stop(bssm_ctx *ctx, float *ms_out);

/* ------------------------------------------------------------------------- *
 * Resamplers.  Replace the three registered .Call routines
 *   _bayesSSM_resample_{multinomial,stratified,systematic}_cpp(n, weights)
 * (src/RcppExports.cpp:15-48, src/resampling.cpp:5-66, R/RcppExports.R:4-14).
 * weights: n doubles (host).  u: uniforms in (0,1) drawn by the caller (the R
 * shim draws them with unif_rand() under GetRNGstate so set.seed() keeps its
 * meaning): n for stratified / multinomial, 1 for systematic.  idx_out: n
 * 1-based ancestor indices.  The cdf reproduces the reference's sequential
 * double-precision cumsum BIT-EXACTLY (parallel exact-rounding scan).
 * Multinomial is the natural-order inverse-CDF draw, distributionally equal
 * to Rcpp::sample but not stream-identical.
 * ------------------------------------------------------------------------- */
int bssm_resample_stratified(bssm_ctx *ctx, int n, const double *weights, const double *u, int32_t *idx_out);
int bssm_resample_systematic(bssm_ctx *ctx, int n, const double *weights, double u, int32_t *idx_out);
int bssm_resample_multinomial(bssm_ctx *ctx, int n, const double *weights, const double *u, int32_t *idx_out);
/* diagnostic: the exact cdf and total the resamplers use (tests) */
int bssm_resample_cdf(bssm_ctx *ctx, int n, const double *weights, double *cdf_out, double *total_out,
                      int64_t *n_serial_out);
/* Device-resident batched variant: `batch` independent weight vectors [batch][n]
 * already in HBM, uniforms [batch][n] (or [batch] for systematic) in HBM,
 * output [batch][n] in HBM.  d_status: [batch] ints (BSSM_OK / error per vector). */
int bssm_resample_device(bssm_ctx *ctx, int resample_fn, int batch, int n, const double *d_weights,
                         const double *d_u, int32_t *d_idx_out, int *d_status);
/* scratch device memory helpers for callers that keep data in HBM (bench / R external pointers) */
int bssm_dev_alloc(bssm_ctx *ctx, size_t bytes, void **d_ptr);
int bssm_dev_free(bssm_ctx *ctx, void *d_ptr);
int bssm_dev_upload(bssm_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int bssm_dev_download(bssm_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
/* fills d_ptr[0..n) with Philox uniforms in (0,1) (same generator as the engine) */
int bssm_dev_fill_uniform(bssm_ctx *ctx, double *d_ptr, size_t n, uint64_t seed);

/* ------------------------------------------------------------------------- *
 * Particle filters.  Replace .particle_filter_core (R/particle_filter_core.R:19-267)
 * as called by bootstrap_filter (R/bootstrap_filter.R:129-171), auxiliary_filter
 * (R/auxiliary_filter.R:163-216) and resample_move_filter
 * (R/resample_move_filter.R:190-236).  One call runs `num_filters` independent
 * filters (same y, per-filter theta) as one batched launch.
 * ------------------------------------------------------------------------- */
typedef struct {
  /* injected noise (parity mode), all double, particle index fastest; NULL => Philox.
   * Shapes as in oracle/pf_oracle.h (orc_noise_buffers); shared by every filter of the batch. */
  const double *z_init, *u_init, *z_trans, *u_trans, *z_trans2, *u_trans2;
  const double *u_resample, *u_resample_aux, *z_move, *u_move;
} bssm_noise_buffers;

typedef struct {
  int model;              /* BSSM_MODEL_* or id from bssm_model_compile */
  int algorithm;          /* BSSM_BPF / APF / RMPF */
  int resample_algorithm; /* BSSM_SIS / SISR / SISAR */
  int resample_fn;        /* BSSM_STRATIFIED / SYSTEMATIC / MULTINOMIAL */
  double threshold;       /* absolute ESS count; < 0 => reference default (R/particle_filter_core.R:44-50) */
  int num_particles;      /* N */
  int num_obs;            /* T = nrow(y) */
  int dy;                 /* ncol(y) */
  const int *obs_times;   /* [T] strictly usable as in R/particle_filter_core.R:71-73; NULL => 1..T */
  int num_filters;        /* batch size C (>= 1) */
  int precision;          /* BSSM_F32 (throughput) or BSSM_F64 (parity) */
  uint64_t seed;          /* Philox key */
  uint32_t run_id;        /* mixed into the key; distinguishes repeated invocations */
  uint32_t stream_base;   /* filter c uses Philox stream stream_base + c */
  const bssm_noise_buffers *noise; /* NULL => Philox */
  int return_particles;   /* fill particles_history / weights_history */
  int exact_resampling;   /* 1: cdf bit-exact vs the sequential reference cumsum; 0: plain parallel fp64 scan;
                             -1 => auto (1 for BSSM_F64, 0 for BSSM_F32) */
  int engine;             /* BSSM_ENGINE_AUTO / GENERAL / PERSISTENT / STREAM */
  int carry_weights;      /* 0: the reference's rule (weights of a step are its likelihoods only, R/particle_filter_core.R:204-209, SURVEY
                             App. A1); 1: DEVIATION, standard SMC: weights carried over steps that do not resample (BPF / RMPF, general
                             kernels) */
} bssm_filter_config;

typedef struct {
  /* all host buffers, caller-allocated; NULL pointers are skipped */
  double *loglike;           /* [C] */
  double *loglike_history;   /* [C][T]   cumulative, R/particle_filter_core.R:208-209 */
  double *ess;               /* [C][T+1] */
  double *state_est;         /* [C][T+1][d] */
  double *particles_history; /* [C][T+1][d][N] (return_particles) */
  double *weights_history;   /* [C][T+1][N]    (return_particles) */
  int32_t *status;           /* [C] BSSM_OK / BSSM_ERR_NAN_WEIGHT ... */
  int32_t *early_exit;       /* [C] 1 if all log-weights < -1e8 (R/particle_filter_core.R:189-202) */
  int32_t *n_resampled;      /* [C] steps where (second-stage) resampling fired */
  int32_t *ancestors_history;     /* [C][T][N] 1-based, 0 where no resampling (test aid; general engine) */
  int32_t *ancestors_aux_history; /* [C][T][N] APF first stage */
  float kernel_ms;           /* device time of the engine launch (CUDA events) */
} bssm_filter_result;

int bssm_model_dims(bssm_ctx *ctx, int model, int *d, int *ntheta, int *nconst);
/* normals / uniforms one particle consumes in init, one transition, one RMPF move (sizes of bssm_noise_buffers) */
int bssm_model_noise_dims(bssm_ctx *ctx, int model, int *nz_init, int *nu_init, int *nz_trans, int *nu_trans,
                          int *nz_move, int *nu_move);
/* theta: [C][ntheta + nconst] host doubles (parameters, then model constants) */
int bssm_filter_run(bssm_ctx *ctx, const bssm_filter_config *cfg, const double *y, const double *theta,
                    bssm_filter_result *res);
/* Device-resident variant: y [T][dy] and theta [C][ntheta+nconst] already in HBM; only
 * d_loglike [C] is written (device).  No host<->device traffic.  Used for bench `value`
 * and by the PMMH driver. */
int bssm_filter_run_device(bssm_ctx *ctx, const bssm_filter_config *cfg, const double *d_y,
                           const double *d_theta, double *d_loglike, float *kernel_ms);

/* ------------------------------------------------------------------------- *
 * Particle-sharded single filter (one filter too large for one GPU, N >= 2^28):
 * the counterpart of running .particle_filter_core (R/particle_filter_core.R:19-267)
 * on one huge particle set.  One process per GPU; rank g keeps a contiguous block
 * of the particles.  Per observation the ranks all-gather one 64-byte record
 * (ncclAllGather over NVLink); the log-normaliser, ESS and resampling decision are
 * derived identically on every rank, the exclusive prefix of the per-rank weight
 * totals gives each rank its cdf offset, and each rank resamples the offspring of
 * its own particles -- no particle crosses NVLink.  Philox streams are keyed by the
 * global particle index: the result equals the one-GPU result up to the summation
 * order of the normaliser.
 *   bssm_shard_unique_id   rank 0: a 128-byte NCCL unique id, to be sent to the other ranks
 *   bssm_shard_init        every rank: join the group (world == 1 needs no id and no NCCL)
 *   bssm_filter_run_sharded  collective; cfg->num_particles is the GLOBAL count, cfg->num_filters
 *                          must be 1, BPF + stratified / systematic, built-in 1-D models.
 *                          capacity_factor (>= 1, default 1.5): storage per rank as a multiple of
 *                          N / world; BSSM_ERR_CAPACITY in status if a rank's share outgrows it.
 * nccl_lib_path: path of libnccl.so.2 (NULL: $BSSM_NCCL_LIB, then the loader's search path).
 * ------------------------------------------------------------------------- */
int bssm_shard_unique_id(const char *nccl_lib_path, void *id_out_128);
int bssm_shard_init(bssm_ctx *ctx, const char *nccl_lib_path, int rank, int world, const void *id_128);
int bssm_shard_finalize(bssm_ctx *ctx);
/* Optional, after bssm_shard_init on every rank: move the per-observation exchange from ncclAllGather into the filter
 * kernel itself.  Each rank exports a 64-byte CUDA IPC handle of its inbox, the host gathers the `world` handles in rank
 * order (the same channel that carried the unique id) and every rank attaches them; from then on the merging block of
 * the propagate / weight kernel stores its record into every rank's inbox over NVLink and polls its own.  Results are
 * identical to the NCCL form (same records, same summation order).  Collective: all ranks attach, or none does.
 * bssm_shard_peer_active: 1 while attached.  A peer that never arrives shows as status BSSM_ERR_NCCL after a timeout. */
int bssm_shard_peer_export(bssm_ctx *ctx, void *handle_out_64);
int bssm_shard_peer_attach(bssm_ctx *ctx, const void *handles_world_x_64);
int bssm_shard_peer_detach(bssm_ctx *ctx);
int bssm_shard_peer_active(const bssm_ctx *ctx);
/* initial block partition (boundaries at multiples of 4): rank's slice [goff, goff + nloc) */
int bssm_shard_partition(int n, int world, int rank, int64_t *goff_out, int *nloc_out);
int bssm_filter_run_sharded(bssm_ctx *ctx, const bssm_filter_config *cfg, const double *y, const double *theta,
                            double capacity_factor, bssm_filter_result *res, int *n_local_final);

/* ------------------------------------------------------------------------- *
 * NVRTC models: CUDA device-function snippets instead of R closures
 * (init_fn / transition_fn / log_likelihood_fn / aux_log_likelihood_fn /
 * move_fn, R/particle_filter-doc.R:11-18; precedent: the cppFunction
 * transition in vignettes/articles/detailed-overview.Rmd:408-466).
 * `src` defines `struct UserModel` (see bayesssm_b200/csrc/bssm_models.cuh for
 * the contract); compiled for compute_100a / sm_100a and cached in the context.
 * ------------------------------------------------------------------------- */
int bssm_model_compile(bssm_ctx *ctx, const char *src, int *model_id_out);
const char *bssm_model_compile_log(bssm_ctx *ctx);

/* ------------------------------------------------------------------------- *
 * PMMH.  Replaces the per-chain closure chain_result (R/pmmh.R:345-505), the
 * pilot chain .run_pilot_chain (R/pmmh_tuning.R:111-317), .pilot_run
 * (R/pmmh_tuning.R:29-64), the transforms (R/utils.R:102-152) and the chain
 * fan-out (R/pmmh.R:511-535).  All chains advance together; proposal,
 * prior, Jacobian, accept/reject and draw storage run on device; no host
 * round trip per iteration.
 * ------------------------------------------------------------------------- */
typedef struct {
  int model, algorithm;      /* pf_wrapper identity */
  int p;                     /* number of parameters (== model ntheta, <= 8) */
  const int *prior_kind; const double *prior_a; const double *prior_b; /* [p] */
  const int *transform;      /* [p] BSSM_TR_* */
  const double *pilot_proposal_sd; /* [p]  (R/pmmh.R:123-126 recycles a scalar) */
  int pilot_n, pilot_m, pilot_reps;               /* R/pmmh.R:33-58 */
  int pilot_resample_algorithm, pilot_resample_fn;
  int m;                     /* draws per chain including burn-in (host drops burn-in) */
  int num_chains;            /* chains handled by THIS call (this rank's shard) */
  uint32_t chain_id_base;    /* global id of this call's first chain (Philox stream) */
  int fixed_num_particles;   /* > 0 overrides target_n (the reference clamps to [50,1000], R/pmmh_tuning.R:54-57) */
  int num_obs, dy; const int *obs_times;
  const double *consts; int nconst;
  int precision;
  uint64_t seed;
  int skip_pilot;            /* 1 => use init_theta as chain start and proposal_chol_in / fixed_num_particles */
  const double *proposal_chol_in; /* [num_chains][p][p] when skip_pilot */
  int engine;                /* BSSM_ENGINE_* for the filter runs */
  int return_latent_state_est; /* R/pmmh.R return_latent_state_est: fill latent_state_chain */
} bssm_pmmh_config;

typedef struct {
  double *pilot_theta_chain;   /* [chains][pilot_m][p] or NULL */
  double *pilot_loglike_chain; /* [chains][pilot_m] or NULL */
  double *pilot_theta_mean;    /* [chains][p] */
  double *pilot_theta_cov;     /* [chains][p][p] */
  double *pilot_loglikes;      /* [chains][pilot_reps] or NULL */
  int32_t *target_n;           /* [chains] */
  double *proposal_chol;       /* [chains][p][p] */
  double *theta_chain;         /* [chains][m][p] */
  double *loglike_chain;       /* [chains][m] */
  int32_t *n_accept;           /* [chains] */
  int32_t *status;             /* [chains] */
  float pilot_ms, main_ms;     /* device time of the two phases */
  double *latent_state_chain;  /* [chains][m][T+1][d] state_est of the filter run behind every draw (R/pmmh.R:420,494-499), or NULL */
  double main_resampled_fraction; /* out: share of the main phase's filter steps that resampled (accepted or not): the measured
                                     r of the roofline's 12 + 28 r bytes per particle-timestep */
} bssm_pmmh_result;

/* init_theta: [num_chains][p] (pilot_init_params) */
int bssm_pmmh_run(bssm_ctx *ctx, const bssm_pmmh_config *cfg, const double *y, const double *init_theta,
                  bssm_pmmh_result *res);

/* MCMC diagnostics of the draws, for all parameters at once: replaces ess() (R/ESS.R:30-104: between/within chain
 * variances, stats::acf of every chain, Geyer's initial monotone sequence) and rhat() (R/rhat.R:27-67: split R-hat, values in
 * [0.99, 1] reported as 1), which the reference calls per parameter on every pmmh() return (R/pmmh.R:570-594).
 * draws: [k][m_total][p] host, chain-major -- the layout of bssm_pmmh_result.theta_chain; with p = 1 it is R's m x k matrix in
 * column-major order.  Iterations [burn_in, m_total) are used (R/pmmh.R:540-545).  ess, rhat: [p] host outputs, either may be
 * NULL; ess needs k >= 2 ("Number of chains must be at least 2."), both need m_total - burn_in >= 2 ("Number of iterations must
 * be at least 2.").  flags: [p] or NULL; bit 0 = a chain has zero variance (ess is NaN: R returns NA with the warning "One or
 * more chains have zero variance."), bit 1 = a half chain has (rhat is NaN, same warning).  device_ms: time of the kernels, or
 * NULL. */
int bssm_mcmc_diagnostics(bssm_ctx *ctx, const double *draws, int k, int m_total, int p, int burn_in, double *ess, double *rhat,
                          int32_t *flags, float *device_ms);

/* transforms, exported so the R side and the tests use one definition (R/utils.R:102-152) */
double bssm_transform(double theta, int tr);
double bssm_back_transform(double z, int tr);
double bssm_log_jacobian(const double *theta, const int *tr, int p);
double bssm_log_prior(int kind, double a, double b, double x);

#ifdef __cplusplus
}
#endif
#endif /* BAYESSSM_B200_H */
