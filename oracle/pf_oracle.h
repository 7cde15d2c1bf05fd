/*
 * pf_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * Plain-C restatement of the bayesSSM hot path (reference: R package
 * BjarkeHautop/bayesSSM 0.7.1.9000).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (bayesssm_b200/) never links, imports or calls it.
 *
 * Parity pinning: R is not installed in this image, so the reference cannot be
 * executed.  The restatement is pinned against every known-answer property the
 * reference's own testthat files hold for this path
 * (tests/testthat/test-resampling.R:48-68,190-202,29-47; test-utils.R:26-59;
 * test-pmmh.R:5-25) and against the committed fixtures under tests/golden/.
 * Bit-level values of loglike / ancestors / draws are NOT pinned by the
 * reference itself ("parity unpinned" for those).
 */
#ifndef PF_ORACLE_H
#define PF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes (mirror the Rcpp::stop strings of src/resampling.cpp:6,8) */
#define ORC_OK 0
#define ORC_ERR_NEGATIVE_WEIGHT 1 /* "Weights must be non-negative" */
#define ORC_ERR_ZERO_SUM 2        /* "Sum of weights must be greater than 0" */
#define ORC_ERR_NAN_WEIGHT 3      /* R: "missing value where TRUE/FALSE needed" */
#define ORC_ERR_BAD_ARG 4
#define ORC_ERR_PRIOR_INIT 5 /* "Initial parameter values are invalid ..." */

/* enums shared (by value) with include/bayesssm_b200.h */
enum { ORC_BPF = 0, ORC_APF = 1, ORC_RMPF = 2 };
enum { ORC_SIS = 0, ORC_SISR = 1, ORC_SISAR = 2 };
enum { ORC_STRATIFIED = 0, ORC_SYSTEMATIC = 1, ORC_MULTINOMIAL = 2, ORC_MULTINOMIAL_SORTED = 3 /* filters only: sorted uniforms, see orc_resample_multinomial_sorted */ };
enum {
  ORC_MODEL_AR_SIN = 0,  /* README.md:137-146 */
  ORC_MODEL_LG = 1,      /* tests/testthat/test-pmmh_tuning.R:163-173 (generalised sigmas) */
  ORC_MODEL_RW_DRIFT = 2,/* tests/testthat/test-auxiliary_filter.R:17-27, test-resample_move_filter.R:17-35 */
  ORC_MODEL_SIR_CB = 3,  /* chain-binomial SIR; obs as stochastic-sir-model.Rmd:306-309 */
  ORC_MODEL_AR_COS = 4,  /* R/pmmh.R:157-159 */
  ORC_MODEL_RW2D = 5,    /* tests/testthat/test-bootstrap_filter.R:211-217 */
  ORC_MODEL_SIR_GILLESPIE = 6 /* the SIR model with the exact daily step of stochastic-sir-model.Rmd:152-176 (epidemic_step); Philox noise only */
};
enum { ORC_PRIOR_FLAT = 0, ORC_PRIOR_NORMAL = 1, ORC_PRIOR_EXP = 2, ORC_PRIOR_UNIF = 3, ORC_PRIOR_HALFNORMAL = 4 };
enum { ORC_TR_IDENTITY = 0, ORC_TR_LOG = 1, ORC_TR_LOGIT = 2 };

/* ---- resamplers: src/resampling.cpp:5-66 (uniforms injected) ---- */
int orc_resample_stratified(int n, const double *w, const double *u, int32_t *idx1);
int orc_resample_systematic(int n, const double *w, double u, int32_t *idx1);
/* multinomial by sorted uniforms from n + 1 exponential spacings (what the streaming engine runs); u: n + 1 uniforms */
int orc_resample_multinomial_sorted(int n, const double *w, const double *u, int32_t *idx1);
/* natural-order inverse-CDF multinomial (what the general kernels are bit-exact against) */
int orc_resample_multinomial_invcdf(int n, const double *w, const double *u, int32_t *idx1);
/* Rcpp::sample(n, n, TRUE, p) restated (Walker alias / sorted inverse CDF), distributional pin only */
int orc_resample_multinomial_rcpp(int n, const double *w, const double *u, int32_t *idx1);
/* intermediate cdf of the stratified/systematic path (for tests) */
int orc_resample_cdf(int n, const double *w, double *cdf, double *total);

/* ---- counter-based noise (restated Philox4x32-10; same keying as the engine) ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_noise_uniform(uint64_t seed, uint32_t run_id, uint32_t stream, uint32_t t, uint32_t tag,
                         uint32_t slot, uint32_t index);
double orc_noise_normal(uint64_t seed, uint32_t run_id, uint32_t stream, uint32_t t, uint32_t tag,
                        uint32_t slot, uint32_t index);

/* ---- injected noise buffers (all double, particle index fastest) ---- */
typedef struct {
  const double *z_init;   /* [nz_init][N] */
  const double *u_init;   /* [nu_init][N] */
  const double *z_trans;  /* [n_time][nz_trans][N], n_time = obs_times[T-1] */
  const double *u_trans;  /* [n_time][nu_trans][N] */
  const double *z_trans2; /* [T][nz_trans][N]  (APF second transition) */
  const double *u_trans2; /* [T][nu_trans][N] */
  const double *u_resample;     /* [T][N] (systematic: element [t][0]) */
  const double *u_resample_aux; /* [T][N] (APF first stage) */
  const double *z_move;   /* [T][nz_move][N] */
  const double *u_move;   /* [T][nu_move][N] */
} orc_noise_buffers;

typedef struct {
  int model;        /* ORC_MODEL_* */
  int algorithm;    /* ORC_BPF / APF / RMPF */
  int resample_algorithm; /* ORC_SIS / SISR / SISAR */
  int resample_fn;  /* ORC_STRATIFIED / ... */
  double threshold; /* absolute ESS count; <0 => reference default (R/particle_filter_core.R:44-50) */
  int num_particles;
  int num_obs;      /* T */
  int dy;           /* columns of y */
  const int *obs_times; /* NULL => 1..T */
  int return_particles;
  /* noise: injected buffers if noise != NULL, else Philox(seed, run_id, stream) */
  const orc_noise_buffers *noise;
  uint64_t seed;
  uint32_t run_id;
  uint32_t stream;
  int carry_weights; /* 0: the reference (R/particle_filter_core.R:204-209: a step's weights are its likelihoods only);
                        1: DEVIATION, weights carried over steps that do not resample (BPF / RMPF) */
} orc_filter_config;

typedef struct {
  double *state_est;       /* [(T+1)][d] row-major (R: (T+1) x d matrix) */
  double *ess;             /* [T+1] */
  double loglike;
  double *loglike_history; /* [T] */
  double *particles_history; /* [(T+1)][d][N] or NULL */
  double *weights_history;   /* [(T+1)][N] or NULL */
  int32_t *ancestors_history; /* [T][N] 1-based, 0 where no resampling; NULL ok (test aid) */
  int32_t *ancestors_aux_history; /* [T][N] APF first stage; NULL ok */
  int early_exit;          /* 1 if R/particle_filter_core.R:189-202 fired */
  int n_resampled;         /* number of steps where second-stage resampling fired */
} orc_filter_result;

int orc_model_dims(int model, int *d, int *ntheta, int *nconst, int *nz_init, int *nu_init,
                   int *nz_trans, int *nu_trans, int *nz_move, int *nu_move);

/* R/particle_filter_core.R:19-267 with R/bootstrap_filter.R, auxiliary_filter.R, resample_move_filter.R */
int orc_particle_filter(const orc_filter_config *cfg, const double *y, const double *theta,
                        orc_filter_result *res);

/* exact Kalman log-likelihood for the LG model (not in the reference) */
double orc_kalman_loglik(int T, const double *y, double phi, double sigma_x, double sigma_y);

/* R/utils.R:102-152 */
double orc_transform(double theta, int tr);
double orc_back_transform(double z, int tr);
double orc_log_jacobian(const double *theta, const int *tr, int p);
double orc_log_prior(int kind, double a, double b, double x);

typedef struct {
  int model, algorithm;  /* pf_wrapper identity */
  int p;                 /* number of parameters (== model ntheta) */
  const int *prior_kind; const double *prior_a; const double *prior_b; /* [p] */
  const int *transform;  /* [p] ORC_TR_* */
  /* tune control R/pmmh.R:33-58 */
  const double *pilot_proposal_sd; /* [p] */
  int pilot_n, pilot_m, pilot_reps;
  int pilot_resample_algorithm, pilot_resample_fn;
  int m, burn_in;
  int fixed_num_particles; /* >0 overrides the [50,1000] clamp (R/pmmh_tuning.R:54-57) */
  int num_obs, dy; const int *obs_times;
  uint64_t seed;
  const double *consts; int nconst; /* model constants appended to theta */
} orc_pmmh_config;

typedef struct {
  double *pilot_theta_chain; /* [pilot_m][p] */
  double *pilot_loglike_chain; /* [pilot_m] */
  double *pilot_theta_mean;  /* [p] */
  double *pilot_theta_cov;   /* [p][p] */
  double *pilot_loglikes;    /* [pilot_reps] */
  int target_n;
  double *proposal_chol;     /* [p][p] lower, of D Sigma D (R/pmmh.R:378-389) */
  double *theta_chain;       /* [m][p] (burn-in NOT removed) */
  double *loglike_chain;     /* [m] */
  int n_accept;
  double *latent_state_chain; /* [m][T+1][d] or NULL: state_est of the filter run behind every draw (R/pmmh.R:420,494-499) */
} orc_pmmh_chain_result;

/* one chain of R/pmmh.R:345-505 (pilot R/pmmh_tuning.R:111-317, pilot run :29-64) */
int orc_pmmh_chain(const orc_pmmh_config *cfg, const double *y, const double *init_theta,
                   uint32_t chain_id, orc_pmmh_chain_result *res);

/* ---- R RNG restatement (Mersenne-Twister + inversion), used only for data simulation
 *      and as the realistic-cost noise source of the CPU baseline ---- */
typedef struct { uint32_t mt[625]; int mti; } orc_rrng;
void orc_rrng_set_seed(orc_rrng *r, uint32_t seed);
double orc_rrng_unif(orc_rrng *r);
double orc_rrng_norm(orc_rrng *r);

/* CPU baseline: bootstrap filter with its own RNG, returns seconds of wall time (R-like cost model) */
double orc_bench_bootstrap_filter(int model, int N, int T, const double *y, const double *theta,
                                  int resample_algorithm, int resample_fn, double threshold,
                                  uint32_t seed, double *loglike_out, int *n_resampled_out);

#ifdef __cplusplus
}
#endif
#endif
