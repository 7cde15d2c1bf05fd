/*
 * pf_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see pf_oracle.h).
 *
 * Every function cites the reference file:line (relative to the bayesSSM
 * source tree) it restates.  Arithmetic follows the reference's evaluation
 * order in IEEE double; base-R `sum()` is restated with a long double
 * accumulator, Rcpp sugar `sum`/`cumsum` with plain
 * sequential double adds (confirmed from the
 * disassembly of src/resampling.o).  Compile with -ffp-contract=off.
 */
#define _POSIX_C_SOURCE 200809L
#include "pf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_LN_SQRT_2PI
#define M_LN_SQRT_2PI 0.918938533204672741780329736406 /* log(sqrt(2*pi)) */
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* Resamplers: src/resampling.cpp                                            */
/* ------------------------------------------------------------------------- */

/* src/resampling.cpp:17-25 / :44-52: validation, total (Rcpp::sum = sequential
 * double adds), prob = w / total (IEEE divide), cum_sum = cumsum(prob)
 * (c[0] = p[0]; c[i] = c[i-1] + p[i]). */
int orc_resample_cdf(int n, const double *w, double *cdf, double *total_out) {
  if (n <= 0) return ORC_ERR_BAD_ARG;
  for (int i = 0; i < n; i++) {
    if (isnan(w[i])) return ORC_ERR_NAN_WEIGHT;
    if (w[i] < 0) return ORC_ERR_NEGATIVE_WEIGHT;
  }
  double total = 0.0;
  for (int i = 0; i < n; i++) total += w[i];
  if (total == 0) return ORC_ERR_ZERO_SUM;
  if (total_out) *total_out = total;
  double run = 0.0;
  for (int i = 0; i < n; i++) {
    double p = w[i] / total;
    run = (i == 0) ? p : run + p;
    cdf[i] = run;
  }
  return ORC_OK;
}

/* src/resampling.cpp:28-37 / :55-63: two-pointer search; first j with
 * cum_sum[j] >= pos (loop runs while cum_sum[j] < pos), clamped to n-1,
 * returned 1-based. */
static void search_positions(int n, const double *cdf, const double *pos, int32_t *idx1) {
  int j = 0;
  for (int i = 0; i < n; i++) {
    while (j < n - 1 && cdf[j] < pos[i]) j++;
    idx1[i] = j + 1;
  }
}

/* src/resampling.cpp:16-40 */
int orc_resample_stratified(int n, const double *w, const double *u, int32_t *idx1) {
  double *cdf = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double *pos = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  int st = orc_resample_cdf(n, w, cdf, NULL);
  if (st == ORC_OK) {
    for (int i = 0; i < n; i++) pos[i] = ((double)i + u[i]) / (double)n; /* :28 */
    search_positions(n, cdf, pos, idx1);
  }
  free(cdf);
  free(pos);
  return st;
}

/* src/resampling.cpp:43-66 */
int orc_resample_systematic(int n, const double *w, double u, int32_t *idx1) {
  double *cdf = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double *pos = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  int st = orc_resample_cdf(n, w, cdf, NULL);
  if (st == ORC_OK) {
    for (int i = 0; i < n; i++) pos[i] = ((double)i + u) / (double)n; /* :55 */
    search_positions(n, cdf, pos, idx1);
  }
  free(cdf);
  free(pos);
  return st;
}

/* Natural-order inverse-CDF multinomial on the same cdf as above: draw i picks
 * the first j with cdf[j] >= u[i] (clamped).  This is the restatement the CUDA
 * multinomial kernel is bit-exact against; it is distributionally identical to
 * src/resampling.cpp:5-13 but does not reproduce Rcpp::sample's index stream. */
int orc_resample_multinomial_invcdf(int n, const double *w, const double *u, int32_t *idx1) {
  double *cdf = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  int st = orc_resample_cdf(n, w, cdf, NULL);
  if (st == ORC_OK) {
    for (int i = 0; i < n; i++) {
      int lo = 0, hi = n - 1; /* first j with cdf[j] >= u, clamp n-1 */
      while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (cdf[mid] < u[i]) lo = mid + 1; else hi = mid;
      }
      idx1[i] = lo + 1;
    }
  }
  free(cdf);
  return st;
}

/* Multinomial resampling by SORTED uniforms: the order statistics of n iid uniforms are the normalised partial sums of n + 1
 * iid Exp(1) spacings, U_(i) = (E_0 + ... + E_i) / (E_0 + ... + E_n), E_k = -log(u[k]) -- so the n draws arrive in increasing
 * order and the two-pointer search of src/resampling.cpp:32-37 serves them like the stratified positions.  Same law as
 * src/resampling.cpp:5-13 (offspring counts ~ Multinomial(n, p)); the ancestors come out sorted, where Rcpp::sample returns
 * them in draw order -- for a particle filter only the multiset matters.  This is what the streaming engine's multinomial
 * path is checked against (u: n + 1 uniforms). */
int orc_resample_multinomial_sorted(int n, const double *w, const double *u, int32_t *idx1) {
  double *cdf = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double *pos = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  int st = orc_resample_cdf(n, w, cdf, NULL);
  if (st == ORC_OK) {
    double total = 0.0;
    for (int i = 0; i <= n; i++) total += -log(u[i]);
    double run = 0.0;
    for (int i = 0; i < n; i++) { run += -log(u[i]); pos[i] = run / total; }
    search_positions(n, cdf, pos, idx1);
  }
  free(cdf);
  free(pos);
  return st;
}

/* src/resampling.cpp:5-13 -> Rcpp::sample(n, n, true, prob).  Rcpp (unpinned,
 * DESCRIPTION:28) is absent from the reference tree; its published algorithm
 * (Rcpp/sugar/functions/sample.h, mirroring R's src/main/random.c) is:
 *   Normalize: p /= sum(p);  nc = #{ n*p[i] > 0.1 };
 *   nc > 200: Walker alias (WalkerSample);  else: ProbSampleReplace
 *   (revsort descending with permutation, cumulate, rU <= p[j]).
 * u[] supplies the unif_rand() stream in consumption order (one per draw). */
typedef struct { double p; int i; } orc_pi;
static int cmp_desc(const void *a, const void *b) {
  const orc_pi *x = (const orc_pi *)a, *y = (const orc_pi *)b;
  if (x->p > y->p) return -1;
  if (x->p < y->p) return 1;
  return (x->i > y->i) - (x->i < y->i); /* tie order of R's revsort is unspecified */
}
int orc_resample_multinomial_rcpp(int n, const double *w, const double *u, int32_t *idx1) {
  if (n <= 0) return ORC_ERR_BAD_ARG;
  for (int i = 0; i < n; i++) {
    if (isnan(w[i])) return ORC_ERR_NAN_WEIGHT;
    if (w[i] < 0) return ORC_ERR_NEGATIVE_WEIGHT;
  }
  double total = 0.0;
  for (int i = 0; i < n; i++) total += w[i];
  if (total == 0) return ORC_ERR_ZERO_SUM;
  double *p = (double *)malloc(sizeof(double) * (size_t)n);
  for (int i = 0; i < n; i++) p[i] = w[i] / total; /* src/resampling.cpp:10 */
  double s = 0.0;                                  /* Normalize (third normalisation) */
  for (int i = 0; i < n; i++) s += p[i];
  for (int i = 0; i < n; i++) p[i] /= s;
  int nc = 0;
  for (int i = 0; i < n; i++) if ((double)n * p[i] > 0.1) nc++;
  if (nc > 200) {
    /* Walker alias */
    int *HL = (int *)malloc(sizeof(int) * (size_t)n);
    int *a = (int *)calloc((size_t)n, sizeof(int));
    double *q = (double *)malloc(sizeof(double) * (size_t)n);
    int *H = HL - 1, *L = HL + n;
    for (int i = 0; i < n; i++) {
      q[i] = p[i] * n;
      if (q[i] < 1.0) *++H = i; else *--L = i;
    }
    if (H >= HL && L < HL + n) {
      for (int k = 0; k < n - 1; k++) {
        int i = HL[k], j = *L;
        a[i] = j;
        q[j] += q[i] - 1;
        if (q[j] < 1.0) L++;
        if (L >= HL + n) break;
      }
    }
    for (int i = 0; i < n; i++) q[i] += i;
    for (int i = 0; i < n; i++) {
      double rU = u[i] * n;
      int k = (int)rU;
      idx1[i] = (rU < q[k]) ? k + 1 : a[k] + 1;
    }
    free(HL); free(a); free(q);
  } else {
    orc_pi *sp = (orc_pi *)malloc(sizeof(orc_pi) * (size_t)n);
    for (int i = 0; i < n; i++) { sp[i].p = p[i]; sp[i].i = i + 1; }
    qsort(sp, (size_t)n, sizeof(orc_pi), cmp_desc);
    for (int i = 1; i < n; i++) sp[i].p += sp[i - 1].p;
    int nm1 = n - 1;
    for (int i = 0; i < n; i++) {
      double rU = u[i];
      int j;
      for (j = 0; j < nm1; j++) if (rU <= sp[j].p) break;
      idx1[i] = sp[j].i;
    }
    free(sp);
  }
  free(p);
  return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* Counter-based noise: Philox4x32-10 (Salmon et al. 2011), restated.        */
/* Keying (shared BY SPECIFICATION with the engine):*/
/*   key = (seed_lo, seed_hi ^ run_id)                                        */
/*   ctr = (index>>2, t, stream, tag | slot<<8)                               */
/*   particle `index` uses word (index&3) for uniforms; Box-Muller pair       */
/*   (index&3)>>1 -> words (2p, 2p+1), cos branch for even index, sin for odd */
/* ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void noise_words(uint64_t seed, uint32_t run_id, uint32_t stream, uint32_t t, uint32_t tag,
                        uint32_t slot, uint32_t index, uint32_t out[4]) {
  uint32_t ctr[4] = {index >> 2, t, stream, tag | (slot << 8)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ run_id};
  orc_philox4x32_10(ctr, key, out);
}
static double word_to_unit(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

double orc_noise_uniform(uint64_t seed, uint32_t run_id, uint32_t stream, uint32_t t, uint32_t tag,
                         uint32_t slot, uint32_t index) {
  uint32_t w[4];
  noise_words(seed, run_id, stream, t, tag, slot, index, w);
  return word_to_unit(w[index & 3]);
}
double orc_noise_normal(uint64_t seed, uint32_t run_id, uint32_t stream, uint32_t t, uint32_t tag,
                        uint32_t slot, uint32_t index) {
  uint32_t w[4];
  noise_words(seed, run_id, stream, t, tag, slot, index, w);
  int pr = (index & 3) >> 1;
  double u1 = word_to_unit(w[2 * pr]), u2 = word_to_unit(w[2 * pr + 1]);
  double r = sqrt(-2.0 * log(u1));
  double ang = 2.0 * M_PI * u2;
  return (index & 1) ? r * sin(ang) : r * cos(ang);
}

/* noise tags */
enum {
  TAG_INIT_Z = 1, TAG_TRANS_Z = 2, TAG_TRANS2_Z = 3, TAG_RESAMP_U = 4, TAG_RESAMP_AUX_U = 5,
  TAG_MOVE_Z = 6, TAG_MOVE_U = 7, TAG_TRANS_U = 8, TAG_TRANS2_U = 9, TAG_INIT_U = 10,
  TAG_TRANS_DYN = 11, TAG_TRANS2_DYN = 12,
  TAG_THETA_Z = 16, TAG_THETA_U = 17
};
#define T_INIT 0xFFFFFFFFu

/* ------------------------------------------------------------------------- */
/* R densities                                        */
/* ------------------------------------------------------------------------- */
static double r_dnorm_log(double x, double mu, double sigma) {
  double z = (x - mu) / sigma;
  return -(M_LN_SQRT_2PI + 0.5 * z * z + log(sigma));
}
static double r_dpois_log(double y, double lambda) {
  if (lambda == 0) return (y == 0) ? 0.0 : -INFINITY;
  return y * log(lambda) - lambda - lgamma(y + 1.0);
}
/* Binomial(n, p) by sequential cdf inversion from ONE uniform (chain-binomial SIR). */
static double binom_inversion(double nd, double p, double u) {
  int n = (int)nd;
  if (n <= 0 || p <= 0) return 0.0;
  if (p >= 1) return (double)n;
  double q = 1.0 - p, r = p / q;
  double pmf = exp((double)n * log1p(-p));
  double cdf = pmf;
  int k = 0;
  while (u > cdf && k < n) {
    k++;
    pmf *= ((double)(n - k + 1) / (double)k) * r;
    cdf += pmf;
  }
  return (double)k;
}

/* ------------------------------------------------------------------------- */
/* Built-in models.  theta = parameters then consts.  */
/* ------------------------------------------------------------------------- */
typedef struct {
  int d, ntheta, nconst, nz_init, nu_init, nz_trans, nu_trans, nz_move, nu_move;
} model_dims;

static int get_dims(int model, model_dims *m) {
  memset(m, 0, sizeof(*m));
  switch (model) {
    case ORC_MODEL_AR_SIN: case ORC_MODEL_LG: case ORC_MODEL_AR_COS:
      m->d = 1; m->ntheta = 3; m->nz_init = 1; m->nz_trans = 1; m->nz_move = 1; m->nu_move = 1; return 0;
    case ORC_MODEL_RW_DRIFT:
      m->d = 1; m->ntheta = 2; m->nz_init = 1; m->nz_trans = 1; m->nz_move = 1; m->nu_move = 1; return 0;
    case ORC_MODEL_SIR_CB:
      m->d = 2; m->ntheta = 2; m->nconst = 2; m->nu_trans = 2; m->nu_move = 2; return 0;
    case ORC_MODEL_SIR_GILLESPIE: /* uniforms on demand (dyn_u), none by slot */
      m->d = 2; m->ntheta = 2; m->nconst = 2; m->nu_move = 2; return 0;
    case ORC_MODEL_RW2D:
      m->d = 2; m->ntheta = 1; m->nz_init = 2; m->nz_trans = 2; return 0;
  }
  return -1;
}
int orc_model_dims(int model, int *d, int *ntheta, int *nconst, int *nz_init, int *nu_init,
                   int *nz_trans, int *nu_trans, int *nz_move, int *nu_move) {
  model_dims m;
  if (get_dims(model, &m)) return ORC_ERR_BAD_ARG;
  *d = m.d; *ntheta = m.ntheta; *nconst = m.nconst; *nz_init = m.nz_init; *nu_init = m.nu_init;
  *nz_trans = m.nz_trans; *nu_trans = m.nu_trans; *nz_move = m.nz_move; *nu_move = m.nu_move;
  return ORC_OK;
}

/* x: d values of ONE particle; z/u: the particle's normals / uniforms by slot */
static void model_init(int model, double *x, const double *th, const double *z, const double *u) {
  (void)u;
  switch (model) {
    case ORC_MODEL_AR_SIN: case ORC_MODEL_LG: case ORC_MODEL_AR_COS: case ORC_MODEL_RW_DRIFT:
      x[0] = 0.0 + 1.0 * z[0]; break; /* rnorm(N, 0, 1): README.md:137-139 */
    case ORC_MODEL_SIR_CB: case ORC_MODEL_SIR_GILLESPIE:
      x[0] = th[2] - th[3]; x[1] = th[3]; break; /* stochastic-sir-model.Rmd:285-292 */
    case ORC_MODEL_RW2D:
      x[0] = z[0]; x[1] = z[1]; break; /* test-bootstrap_filter.R:211 */
  }
}
/* uniforms on demand for one particle's transition (engine: DynU, bssm_common.cuh): uniform k is word k & 3 of
   Philox(ctr = (particle, t, stream, tag | (k >> 2) << 8)) */
typedef struct { uint64_t seed; uint32_t run_id, stream, t, tag, particle; } dyn_u;
static double dyn_uniform(const dyn_u *du, int k) {
  uint32_t ctr[4] = {du->particle, du->t, du->stream, du->tag | ((uint32_t)(k >> 2) << 8)};
  uint32_t key[2] = {(uint32_t)du->seed, (uint32_t)(du->seed >> 32) ^ du->run_id}, w[4];
  orc_philox4x32_10(ctr, key, w);
  return word_to_unit(w[k & 3]);
}
static void model_transition(int model, double *x, const double *th, int t, const double *z,
                             const double *u, const dyn_u *du) {
  (void)t;
  switch (model) {
    case ORC_MODEL_SIR_GILLESPIE: { /* stochastic-sir-model.Rmd:152-176: epidemic_step(state, lambda, gamma, n_total) */
      double s = x[0], i = x[1], tt = 0.0;
      const double lam = th[0] / th[2], gam = th[1];
      const int max_events = 2 * (int)th[2] + 8;
      for (int e = 0; e < max_events && i > 0.0; e++) {
        const double rate_inf = lam * s * i, rate_rem = gam * i, rate = rate_inf + rate_rem;
        if (!(rate > 0.0)) break;
        const double dt = -log(dyn_uniform(du, 2 * e)) / rate; /* rexp(1, rate_total) */
        if (tt + dt > 1.0) break;
        tt += dt;
        if (dyn_uniform(du, 2 * e + 1) < rate_inf / rate) { s -= 1.0; i += 1.0; } else { i -= 1.0; }
      }
      x[0] = s; x[1] = i; break;
    }
    case ORC_MODEL_AR_SIN: case ORC_MODEL_AR_COS: /* README.md:140-143 */
      x[0] = th[0] * x[0] + sin(x[0]) + (0.0 + th[1] * z[0]); break;
    case ORC_MODEL_LG: /* test-pmmh_tuning.R:166-168 */
      x[0] = th[0] * x[0] + (0.0 + th[1] * z[0]); break;
    case ORC_MODEL_RW_DRIFT: /* test-auxiliary_filter.R:18-20 */
      x[0] = x[0] + (th[0] + 1.0 * z[0]); break;
    case ORC_MODEL_SIR_CB: {
      double S = x[0], I = x[1], pop = th[2];
      if (I == 0) break; /* stochastic-sir-model.Rmd:296-298 */
      double p_inf = 1.0 - exp(-th[0] * I / pop);
      double p_rec = 1.0 - exp(-th[1]);
      double ni = binom_inversion(S, p_inf, u[0]);
      double nr = binom_inversion(I, p_rec, u[1]);
      x[0] = S - ni; x[1] = I + ni - nr; break;
    }
    case ORC_MODEL_RW2D: /* test-pmmh.R:625-627 */
      x[0] = x[0] + (th[0] + z[0]); x[1] = x[1] + (th[0] + z[1]); break;
  }
}
static double model_loglik(int model, const double *y, const double *x, const double *th, int t) {
  (void)t;
  switch (model) {
    case ORC_MODEL_AR_SIN: case ORC_MODEL_LG: return r_dnorm_log(y[0], x[0], th[2]); /* README.md:144-146 */
    case ORC_MODEL_AR_COS: return r_dnorm_log(y[0], cos(x[0]), th[2]); /* R/pmmh.R:157-159 */
    case ORC_MODEL_RW_DRIFT: return r_dnorm_log(y[0], x[0], th[1]);
    case ORC_MODEL_SIR_CB: case ORC_MODEL_SIR_GILLESPIE: return r_dpois_log(y[0], x[1]); /* stochastic-sir-model.Rmd:306-309 */
    case ORC_MODEL_RW2D: return 1.0; /* test-bootstrap_filter.R:215 */
  }
  return NAN;
}
static double model_aux_loglik(int model, const double *y, const double *x, const double *th, int t) {
  switch (model) {
    case ORC_MODEL_RW_DRIFT: return r_dnorm_log(y[0], x[0] + th[0], th[1]); /* test-auxiliary_filter.R:24-27 */
    case ORC_MODEL_SIR_CB: case ORC_MODEL_SIR_GILLESPIE: {
      double S = x[0], I = x[1], pop = th[2];
      double p_inf = 1.0 - exp(-th[0] * I / pop), p_rec = 1.0 - exp(-th[1]);
      return r_dpois_log(y[0], I + S * p_inf - I * p_rec);
    }
    case ORC_MODEL_AR_SIN: case ORC_MODEL_AR_COS: { /* forecast mean */
      double f = th[0] * x[0] + sin(x[0]);
      return model == ORC_MODEL_AR_SIN ? r_dnorm_log(y[0], f, th[2]) : r_dnorm_log(y[0], cos(f), th[2]);
    }
    case ORC_MODEL_LG: return r_dnorm_log(y[0], th[0] * x[0], th[2]);
    default: return model_loglik(model, y, x, th, t);
  }
}
static void model_move(int model, double *x, const double *y, const double *th, int t, const double *z,
                       const double *u) {
  switch (model) {
    case ORC_MODEL_SIR_CB: case ORC_MODEL_SIR_GILLESPIE: {
      double prop[2] = {x[0], x[1] + (u[0] < 0.5 ? -1.0 : 1.0)};
      if (prop[1] < 0 || prop[1] > th[2] - x[0]) return;
      double lc = model_loglik(model, y, x, th, t), lp = model_loglik(model, y, prop, th, t);
      if (log(u[1]) < lp - lc) x[1] = prop[1];
      return;
    }
    case ORC_MODEL_RW2D: return;
    default: { /* test-resample_move_filter.R:24-35: RW(0.1) MH on g */
      double prop = x[0] + (0.0 + 0.1 * z[0]);
      double lc = model_loglik(model, y, x, th, t), lp = model_loglik(model, y, &prop, th, t);
      if (log(u[0]) < lp - lc) x[0] = prop;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* Filter core: R/particle_filter_core.R:19-267                              */
/* ------------------------------------------------------------------------- */
static double r_sum(const double *x, int n) { /* base R sum(): long double accumulator */
  long double s = 0.0L;
  for (int i = 0; i < n; i++) s += x[i];
  return (double)s;
}

typedef struct {
  const orc_filter_config *cfg;
  model_dims md;
  int N;
} fctx;

static double get_z(const fctx *f, const double *buf, int nslot, uint32_t tag, uint32_t t, int row,
                    int slot, int i) {
  if (f->cfg->noise) return buf[((size_t)row * nslot + slot) * f->N + i];
  return orc_noise_normal(f->cfg->seed, f->cfg->run_id, f->cfg->stream, t, tag, (uint32_t)slot, (uint32_t)i);
}
static double get_u(const fctx *f, const double *buf, int nslot, uint32_t tag, uint32_t t, int row,
                    int slot, int i) {
  if (f->cfg->noise) return buf[((size_t)row * nslot + slot) * f->N + i];
  return orc_noise_uniform(f->cfg->seed, f->cfg->run_id, f->cfg->stream, t, tag, (uint32_t)slot, (uint32_t)i);
}

/* particles are stored [d][N] (R: N x d column-major matrix) */
static void do_transition(const fctx *f, double *px, const double *th, int tnow, int second, int obs_i) {
  const orc_noise_buffers *nb = f->cfg->noise;
  int N = f->N, d = f->md.d;
  double x[4], z[4], u[4];
  for (int i = 0; i < N; i++) {
    for (int k = 0; k < d; k++) x[k] = px[(size_t)k * N + i];
    if (!second) {
      for (int s = 0; s < f->md.nz_trans; s++)
        z[s] = get_z(f, nb ? nb->z_trans : NULL, f->md.nz_trans, TAG_TRANS_Z, (uint32_t)(tnow - 1), tnow - 1, s, i);
      for (int s = 0; s < f->md.nu_trans; s++)
        u[s] = get_u(f, nb ? nb->u_trans : NULL, f->md.nu_trans, TAG_TRANS_U, (uint32_t)(tnow - 1), tnow - 1, s, i);
    } else {
      for (int s = 0; s < f->md.nz_trans; s++)
        z[s] = get_z(f, nb ? nb->z_trans2 : NULL, f->md.nz_trans, TAG_TRANS2_Z, (uint32_t)obs_i, obs_i, s, i);
      for (int s = 0; s < f->md.nu_trans; s++)
        u[s] = get_u(f, nb ? nb->u_trans2 : NULL, f->md.nu_trans, TAG_TRANS2_U, (uint32_t)obs_i, obs_i, s, i);
    }
    dyn_u du = {f->cfg->seed, f->cfg->run_id, f->cfg->stream, second ? (uint32_t)obs_i : (uint32_t)(tnow - 1),
                second ? TAG_TRANS2_DYN : TAG_TRANS_DYN, (uint32_t)i};
    model_transition(f->cfg->model, x, th, tnow, z, u, &du);
    for (int k = 0; k < d; k++) px[(size_t)k * N + i] = x[k];
  }
}

/* R/resampling.R:13-69: ancestors from the C++ resampler, then gather rows */
static int do_resample(const fctx *f, const double *weights, int obs_i, int aux, int32_t *anc1) {
  const orc_noise_buffers *nb = f->cfg->noise;
  int N = f->N;
  const double *buf = nb ? (aux ? nb->u_resample_aux : nb->u_resample) : NULL;
  uint32_t tag = aux ? TAG_RESAMP_AUX_U : TAG_RESAMP_U;
  int st;
  if (f->cfg->resample_fn == ORC_SYSTEMATIC) {
    double u = get_u(f, buf, 1, tag, (uint32_t)obs_i, obs_i, 0, 0);
    st = orc_resample_systematic(N, weights, u, anc1);
  } else {
    /* ORC_MULTINOMIAL_SORTED consumes one uniform more (n + 1 spacings): Philox mode only (the injected buffers hold N per step) */
    const int sorted = f->cfg->resample_fn == ORC_MULTINOMIAL_SORTED;
    if (sorted && buf) return ORC_ERR_BAD_ARG;
    const int nu = sorted ? N + 1 : N;
    double *u = (double *)malloc(sizeof(double) * (size_t)nu);
    for (int i = 0; i < nu; i++) u[i] = get_u(f, buf, 1, tag, (uint32_t)obs_i, obs_i, 0, i);
    st = (f->cfg->resample_fn == ORC_STRATIFIED) ? orc_resample_stratified(N, weights, u, anc1)
         : (sorted ? orc_resample_multinomial_sorted(N, weights, u, anc1) : orc_resample_multinomial_invcdf(N, weights, u, anc1));
    free(u);
  }
  return st;
}
static void gather(int N, int d, double *px, const int32_t *anc1, double *tmp) {
  for (int k = 0; k < d; k++) {
    for (int i = 0; i < N; i++) tmp[i] = px[(size_t)k * N + (anc1[i] - 1)];
    memcpy(px + (size_t)k * N, tmp, sizeof(double) * (size_t)N);
  }
}

int orc_particle_filter(const orc_filter_config *cfg, const double *y, const double *theta,
                        orc_filter_result *res) {
  fctx f;
  f.cfg = cfg;
  if (get_dims(cfg->model, &f.md)) return ORC_ERR_BAD_ARG;
  const int N = cfg->num_particles, T = cfg->num_obs, d = f.md.d, dy = cfg->dy;
  if (N <= 0 || T < 0 || dy <= 0) return ORC_ERR_BAD_ARG;
  f.N = N;
  const orc_noise_buffers *nb = cfg->noise;
  if (nb && cfg->model == ORC_MODEL_SIR_GILLESPIE) return ORC_ERR_BAD_ARG; /* uniforms on demand: Philox noise only */
  if (cfg->carry_weights && cfg->algorithm == ORC_APF) return ORC_ERR_BAD_ARG;
  int resampled_prev = 1;
  /* R/resample_move_filter.R:228-230: RMPF forces SISR */
  int ralg = (cfg->algorithm == ORC_RMPF) ? ORC_SISR : cfg->resample_algorithm;
  /* R/particle_filter_core.R:44-50 */
  double threshold = cfg->threshold;
  if (threshold < 0 || cfg->algorithm == ORC_RMPF)
    threshold = (ralg == ORC_SIS) ? INFINITY : (ralg == ORC_SISR ? (double)N : (double)N / 2.0);

  double *px = (double *)malloc(sizeof(double) * (size_t)N * d);
  double *lw = (double *)malloc(sizeof(double) * (size_t)N);
  double *auxlw = (double *)malloc(sizeof(double) * (size_t)N);
  double *w = (double *)malloc(sizeof(double) * (size_t)N);
  double *tmp = (double *)malloc(sizeof(double) * (size_t)N);
  int32_t *anc = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
  int status = ORC_OK;

  memset(res->state_est, 0, sizeof(double) * (size_t)(T + 1) * d);
  memset(res->ess, 0, sizeof(double) * (size_t)(T + 1));
  if (res->loglike_history) memset(res->loglike_history, 0, sizeof(double) * (size_t)T);
  if (res->ancestors_history) memset(res->ancestors_history, 0, sizeof(int32_t) * (size_t)T * N);
  if (res->ancestors_aux_history) memset(res->ancestors_aux_history, 0, sizeof(int32_t) * (size_t)T * N);
  res->early_exit = 0;
  res->n_resampled = 0;

  /* :76-88 init_fn */
  {
    double x[4], z[4], u[4];
    for (int i = 0; i < N; i++) {
      for (int s = 0; s < f.md.nz_init; s++) z[s] = get_z(&f, nb ? nb->z_init : NULL, f.md.nz_init, TAG_INIT_Z, T_INIT, 0, s, i);
      for (int s = 0; s < f.md.nu_init; s++) u[s] = get_u(&f, nb ? nb->u_init : NULL, f.md.nu_init, TAG_INIT_U, T_INIT, 0, s, i);
      model_init(cfg->model, x, theta, z, u);
      for (int k = 0; k < d; k++) px[(size_t)k * N + i] = x[k];
    }
  }
  /* :106-116 */
  for (int i = 0; i < N; i++) w[i] = 1.0 / (double)N;
  {
    for (int i = 0; i < N; i++) tmp[i] = w[i] * w[i];
    res->ess[0] = 1.0 / r_sum(tmp, N);
    for (int k = 0; k < d; k++) {
      for (int i = 0; i < N; i++) tmp[i] = px[(size_t)k * N + i] * w[i];
      res->state_est[k] = r_sum(tmp, N);
    }
    if (res->particles_history) memcpy(res->particles_history, px, sizeof(double) * (size_t)N * d);
    if (res->weights_history) memcpy(res->weights_history, w, sizeof(double) * (size_t)N);
  }

  double loglike = 0.0;
  int prev_t = 0;
  for (int i = 0; i < T; i++) { /* :123 */
    int ot = cfg->obs_times ? cfg->obs_times[i] : i + 1;
    int gap = ot - prev_t;
    for (int step = 1; step <= gap; step++) do_transition(&f, px, theta, prev_t + step, 0, i); /* :125-136 */
    prev_t = ot;
    const double *yi = y + (size_t)i * dy;
    double x[4];

    if (cfg->algorithm == ORC_APF) { /* :140-175 */
      for (int j = 0; j < N; j++) {
        for (int k = 0; k < d; k++) x[k] = px[(size_t)k * N + j];
        auxlw[j] = model_aux_loglik(cfg->model, yi, x, theta, prev_t);
        if (isnan(auxlw[j])) { status = ORC_ERR_NAN_WEIGHT; goto done; }
      }
      double max_aux = -INFINITY;
      for (int j = 0; j < N; j++) if (auxlw[j] > max_aux) max_aux = auxlw[j];
      for (int j = 0; j < N; j++) tmp[j] = exp(auxlw[j] - max_aux);
      double s = r_sum(tmp, N);
      for (int j = 0; j < N; j++) w[j] = tmp[j] / s;
      status = do_resample(&f, w, i, 1, anc); /* :155 */
      if (status) goto done;
      if (res->ancestors_aux_history) memcpy(res->ancestors_aux_history + (size_t)i * N, anc, sizeof(int32_t) * (size_t)N);
      gather(N, d, px, anc, tmp);             /* :157 */
      do_transition(&f, px, theta, prev_t, 1, i); /* :159 second transition at the same t */
      for (int j = 0; j < N; j++) {
        for (int k = 0; k < d; k++) x[k] = px[(size_t)k * N + j];
        lw[j] = model_loglik(cfg->model, yi, x, theta, prev_t) - auxlw[anc[j] - 1]; /* :169-175 */
      }
    } else {
      for (int j = 0; j < N; j++) { /* :177-183 */
        for (int k = 0; k < d; k++) x[k] = px[(size_t)k * N + j];
        lw[j] = model_loglik(cfg->model, yi, x, theta, prev_t);
        /* carried-weights mode (not the reference): + log(N w_{t-1}); w is 1/N at the first observation and after a resampling */
        if (cfg->carry_weights && i > 0 && !resampled_prev) lw[j] = lw[j] + log((double)N * w[j]);
      }
    }
    for (int j = 0; j < N; j++) if (isnan(lw[j])) { status = ORC_ERR_NAN_WEIGHT; goto done; }

    /* :189-202 early exit */
    int all_small = 1;
    for (int j = 0; j < N; j++) if (!(lw[j] < -1e8)) { all_small = 0; break; }
    if (all_small) {
      loglike = -INFINITY;
      if (res->loglike_history) res->loglike_history[i] = -INFINITY;
      res->early_exit = 1;
      goto done;
    }
    /* :204-209 */
    double max_logw = -INFINITY;
    for (int j = 0; j < N; j++) if (lw[j] > max_logw) max_logw = lw[j];
    for (int j = 0; j < N; j++) tmp[j] = exp(lw[j] - max_logw);
    double weight_sum = r_sum(tmp, N);
    for (int j = 0; j < N; j++) w[j] = tmp[j] / weight_sum;
    loglike = loglike + (max_logw + log(weight_sum) - log((double)N));
    if (res->loglike_history) res->loglike_history[i] = loglike;
    /* :211-212 */
    for (int j = 0; j < N; j++) tmp[j] = w[j] * w[j];
    double ess = 1.0 / r_sum(tmp, N);
    res->ess[i + 1] = ess;
    /* :214-224 */
    int should = (ralg == ORC_SIS) ? 0 : (ralg == ORC_SISR ? 1 : (ess < threshold));
    resampled_prev = (cfg->algorithm == ORC_RMPF || should);
    if (cfg->algorithm == ORC_RMPF || should) {
      status = do_resample(&f, w, i, 0, anc);
      if (status) goto done;
      if (res->ancestors_history) memcpy(res->ancestors_history + (size_t)i * N, anc, sizeof(int32_t) * (size_t)N);
      gather(N, d, px, anc, tmp);
      for (int j = 0; j < N; j++) w[j] = 1.0 / (double)N;
      res->ess[i + 1] = (double)N;
      res->n_resampled++;
    }
    /* :226-234 */
    if (cfg->algorithm == ORC_RMPF) {
      double z[4], u[4];
      for (int j = 0; j < N; j++) {
        for (int k = 0; k < d; k++) x[k] = px[(size_t)k * N + j];
        for (int s = 0; s < f.md.nz_move; s++) z[s] = get_z(&f, nb ? nb->z_move : NULL, f.md.nz_move, TAG_MOVE_Z, (uint32_t)i, i, s, j);
        for (int s = 0; s < f.md.nu_move; s++) u[s] = get_u(&f, nb ? nb->u_move : NULL, f.md.nu_move, TAG_MOVE_U, (uint32_t)i, i, s, j);
        model_move(cfg->model, x, yi, theta, prev_t, z, u);
        for (int k = 0; k < d; k++) px[(size_t)k * N + j] = x[k];
      }
    }
    /* :237-245 */
    for (int k = 0; k < d; k++) {
      for (int j = 0; j < N; j++) tmp[j] = px[(size_t)k * N + j] * w[j];
      res->state_est[(size_t)(i + 1) * d + k] = r_sum(tmp, N);
    }
    if (res->particles_history) memcpy(res->particles_history + (size_t)(i + 1) * N * d, px, sizeof(double) * (size_t)N * d);
    if (res->weights_history) memcpy(res->weights_history + (size_t)(i + 1) * N, w, sizeof(double) * (size_t)N);
  }
done:
  res->loglike = loglike;
  free(px); free(lw); free(auxlw); free(w); free(tmp); free(anc);
  return status;
}

/* exact Kalman log-likelihood, LG model x0~N(0,1) (SURVEY.md 8c "additional oracle content") */
double orc_kalman_loglik(int T, const double *y, double phi, double sigma_x, double sigma_y) {
  double m = 0.0, P = 1.0, ll = 0.0;
  for (int t = 0; t < T; t++) {
    double mp = phi * m, Pp = phi * phi * P + sigma_x * sigma_x;
    double S = Pp + sigma_y * sigma_y, r = y[t] - mp;
    ll -= 0.5 * (log(2.0 * M_PI) + log(S) + r * r / S);
    double K = Pp / S;
    m = mp + K * r;
    P = (1.0 - K) * Pp;
  }
  return ll;
}

/* ------------------------------------------------------------------------- */
/* Transforms and priors: R/utils.R:102-152, SURVEY.md Appendix F            */
/* ------------------------------------------------------------------------- */
double orc_transform(double th, int tr) { /* R/utils.R:102-112 */
  if (tr == ORC_TR_LOG) return log(th);
  if (tr == ORC_TR_LOGIT) return log(th / (1.0 - th));
  return th;
}
double orc_back_transform(double z, int tr) { /* R/utils.R:122-132 */
  if (tr == ORC_TR_LOG) return exp(z);
  if (tr == ORC_TR_LOGIT) return 1.0 / (1.0 + exp(-z));
  return z;
}
double orc_log_jacobian(const double *th, const int *tr, int p) { /* R/utils.R:142-152 (sic, A14) */
  double s = 0.0;
  for (int j = 0; j < p; j++) {
    if (tr[j] == ORC_TR_LOG) s += log(th[j]);
    else if (tr[j] == ORC_TR_LOGIT) s += log(1.0 / (th[j] * (1.0 - th[j])));
  }
  return s;
}
double orc_log_prior(int kind, double a, double b, double x) {
  switch (kind) {
    case ORC_PRIOR_FLAT: return 0.0;
    case ORC_PRIOR_NORMAL: return r_dnorm_log(x, a, b);
    case ORC_PRIOR_EXP: return (x < 0) ? -INFINITY : log(a) - a * x;
    case ORC_PRIOR_UNIF: return (a <= x && x <= b) ? -log(b - a) : -INFINITY;
    case ORC_PRIOR_HALFNORMAL: return (x < 0) ? -INFINITY : log(2.0) + r_dnorm_log(x, 0.0, a);
  }
  return NAN;
}

/* ------------------------------------------------------------------------- */
/* PMMH: R/pmmh.R:345-505, R/pmmh_tuning.R:29-64,111-317                     */
/* theta-level draws: Philox ctr=(j>>2 word j&3 .., iteration, chain, tag|attempt<<8), */
/* key=(seed_lo, seed_hi ^ phase<<28).  Filter run_id = phase<<28 | iteration.          */
/* ------------------------------------------------------------------------- */
enum { PH_PILOT = 1, PH_PILOT_RUN = 2, PH_MAIN = 3 };

static int run_filter_se(const orc_pmmh_config *c, const double *y, const double *theta, int N, int ralg,
                         int rfn, uint32_t run_id, uint32_t chain_id, double *loglike, double *state_est_out /* [(T+1) d] or NULL */) {
  model_dims md;
  get_dims(c->model, &md);
  orc_filter_config fc;
  memset(&fc, 0, sizeof(fc));
  fc.model = c->model; fc.algorithm = c->algorithm;
  fc.resample_algorithm = ralg; fc.resample_fn = rfn; fc.threshold = -1.0;
  fc.num_particles = N; fc.num_obs = c->num_obs; fc.dy = c->dy; fc.obs_times = c->obs_times;
  fc.noise = NULL; fc.seed = c->seed; fc.run_id = run_id; fc.stream = chain_id;
  double th[16];
  for (int j = 0; j < c->p; j++) th[j] = theta[j];
  for (int j = 0; j < c->nconst; j++) th[c->p + j] = c->consts[j];
  orc_filter_result r;
  memset(&r, 0, sizeof(r));
  r.state_est = (double *)malloc(sizeof(double) * (size_t)(c->num_obs + 1) * md.d);
  r.ess = (double *)malloc(sizeof(double) * (size_t)(c->num_obs + 1));
  int st = orc_particle_filter(&fc, y, th, &r);
  *loglike = r.loglike;
  if (state_est_out) memcpy(state_est_out, r.state_est, sizeof(double) * (size_t)(c->num_obs + 1) * md.d);
  free(r.state_est); free(r.ess);
  return st;
}
static int run_filter(const orc_pmmh_config *c, const double *y, const double *theta, int N, int ralg,
                      int rfn, uint32_t run_id, uint32_t chain_id, double *loglike) {
  return run_filter_se(c, y, theta, N, ralg, rfn, run_id, chain_id, loglike, NULL);
}
static int priors_finite(const orc_pmmh_config *c, const double *th, double *sum_out) {
  double s = 0.0; int ok = 1;
  for (int j = 0; j < c->p; j++) {
    double lp = orc_log_prior(c->prior_kind[j], c->prior_a[j], c->prior_b[j], th[j]);
    if (!isfinite(lp)) ok = 0;
    s += lp;
  }
  *sum_out = s;
  return ok;
}

int orc_pmmh_chain(const orc_pmmh_config *c, const double *y, const double *init_theta,
                   uint32_t chain_id, orc_pmmh_chain_result *res) {
  const int p = c->p;
  if (p < 1 || p > 8) return ORC_ERR_BAD_ARG;
  double cur[8], prop[8], zc[8], zp[8];
  double sum_prior;
  /* R/pmmh_tuning.R:135-143 */
  memcpy(cur, init_theta, sizeof(double) * (size_t)p);
  if (!priors_finite(c, cur, &sum_prior)) return ORC_ERR_PRIOR_INIT;
  double cur_ll;
  int st = run_filter(c, y, cur, c->pilot_n, c->pilot_resample_algorithm, c->pilot_resample_fn,
                      ((uint32_t)PH_PILOT << 28) | 0u, chain_id, &cur_ll); /* :151-165 */
  if (st) return st;
  memcpy(res->pilot_theta_chain, cur, sizeof(double) * (size_t)p);
  res->pilot_loglike_chain[0] = cur_ll;
  const uint64_t seed = c->seed;
  for (int it = 1; it < c->pilot_m; it++) { /* :191-257 */
    double lp_prop_sum = 0.0;
    for (uint32_t attempt = 0;; attempt++) { /* :193-208 re-propose until priors finite */
      if (attempt > 0xFFFFu) return ORC_ERR_BAD_ARG;
      for (int j = 0; j < p; j++) {
        zc[j] = orc_transform(cur[j], c->transform[j]);
        zp[j] = zc[j] + (0.0 + c->pilot_proposal_sd[j] *
                orc_noise_normal(seed, (uint32_t)PH_PILOT << 28, chain_id, (uint32_t)it, TAG_THETA_Z, attempt, (uint32_t)j));
        prop[j] = orc_back_transform(zp[j], c->transform[j]);
      }
      if (priors_finite(c, prop, &lp_prop_sum)) break;
    }
    double lp_cur_sum;
    priors_finite(c, cur, &lp_cur_sum);
    double prop_ll;
    st = run_filter(c, y, prop, c->pilot_n, c->pilot_resample_algorithm, c->pilot_resample_fn,
                    ((uint32_t)PH_PILOT << 28) | (uint32_t)it, chain_id, &prop_ll);
    if (st) return st;
    double num = lp_prop_sum + prop_ll + orc_log_jacobian(prop, c->transform, p); /* :244-246 */
    double den = lp_cur_sum + cur_ll + orc_log_jacobian(cur, c->transform, p);
    double ratio = num - den;
    if (isnan(ratio)) ratio = -INFINITY;
    double u = orc_noise_uniform(seed, (uint32_t)PH_PILOT << 28, chain_id, (uint32_t)it, TAG_THETA_U, 0, 0);
    if (log(u) < ratio) { memcpy(cur, prop, sizeof(double) * (size_t)p); cur_ll = prop_ll; }
    memcpy(res->pilot_theta_chain + (size_t)it * p, cur, sizeof(double) * (size_t)p);
    res->pilot_loglike_chain[it] = cur_ll;
  }
  /* :260-267 mean / cov of the second half, original scale, unbiased */
  int b0 = c->pilot_m / 2, nn = c->pilot_m - b0;
  for (int j = 0; j < p; j++) {
    double s = 0.0;
    for (int it = b0; it < c->pilot_m; it++) s += res->pilot_theta_chain[(size_t)it * p + j];
    res->pilot_theta_mean[j] = s / (double)nn;
  }
  for (int a = 0; a < p; a++)
    for (int b = 0; b < p; b++) {
      double s = 0.0;
      for (int it = b0; it < c->pilot_m; it++)
        s += (res->pilot_theta_chain[(size_t)it * p + a] - res->pilot_theta_mean[a]) *
             (res->pilot_theta_chain[(size_t)it * p + b] - res->pilot_theta_mean[b]);
      res->pilot_theta_cov[a * p + b] = s / (double)(nn - 1);
    }
  /* .pilot_run R/pmmh_tuning.R:29-64.  resample_algorithm / resample_fn are tune_control's pilot settings here too: pmmh() passes
   * them to .run_pilot_chain (R/pmmh.R:366-367), whose `...` is spliced into do.call(.pilot_run, ...) (R/pmmh_tuning.R:292-305);
   * .pilot_run forwards resample_fn and its own `...` to pf_wrapper (:34-50).  run_filter() applies RMPF's SISR override. */
  {
    double mean_ll = 0.0;
    for (int r = 0; r < c->pilot_reps; r++) {
      st = run_filter(c, y, res->pilot_theta_mean, c->pilot_n, c->pilot_resample_algorithm, c->pilot_resample_fn,
                      ((uint32_t)PH_PILOT_RUN << 28) | (uint32_t)r, chain_id, &res->pilot_loglikes[r]);
      if (st) return st;
      mean_ll += res->pilot_loglikes[r];
    }
    mean_ll /= (double)c->pilot_reps;
    double v = 0.0;
    for (int r = 0; r < c->pilot_reps; r++) v += (res->pilot_loglikes[r] - mean_ll) * (res->pilot_loglikes[r] - mean_ll);
    v /= (double)(c->pilot_reps - 1);
    double tn = ceil((double)c->pilot_n * v);
    if (!(tn >= 50)) tn = 50; /* max(target_n, 50); NaN -> 50 is our choice, R would propagate NA */
    if (tn > 1000) tn = 1000;
    res->target_n = (c->fixed_num_particles > 0) ? c->fixed_num_particles : (int)tn;
  }
  /* R/pmmh.R:378-389: D Sigma D, D = diag(dz/dtheta at pilot mean); factor = lower Cholesky
   * (MASS::mvrnorm uses an eigen factor; any L with L L' = Sigma is distributionally identical) */
  {
    double sc[8], S[64];
    for (int j = 0; j < p; j++) {
      double th = res->pilot_theta_mean[j];
      sc[j] = c->transform[j] == ORC_TR_LOG ? 1.0 / th : (c->transform[j] == ORC_TR_LOGIT ? 1.0 / (th * (1.0 - th)) : 1.0);
    }
    for (int a = 0; a < p; a++) for (int b = 0; b < p; b++) S[a * p + b] = sc[a] * res->pilot_theta_cov[a * p + b] * sc[b];
    double *L = res->proposal_chol;
    memset(L, 0, sizeof(double) * (size_t)p * p);
    for (int j = 0; j < p; j++) {
      double dsum = S[j * p + j];
      for (int k = 0; k < j; k++) dsum -= L[j * p + k] * L[j * p + k];
      if (!(dsum > 0)) { for (int i = j; i < p; i++) L[i * p + j] = 0.0; continue; }
      double dj = sqrt(dsum);
      L[j * p + j] = dj;
      for (int i = j + 1; i < p; i++) {
        double s = S[i * p + j];
        for (int k = 0; k < j; k++) s -= L[i * p + k] * L[j * p + k];
        L[i * p + j] = s / dj;
      }
    }
  }
  /* main chain R/pmmh.R:395-500 */
  memcpy(cur, res->pilot_theta_mean, sizeof(double) * (size_t)p);
  /* latent state estimates (R/pmmh.R:400,420,494-499): current_state_est travels with the chain */
  model_dims md_se;
  get_dims(c->model, &md_se);
  const size_t se_len = (size_t)(c->num_obs + 1) * md_se.d;
  double *cur_se = NULL, *prop_se = NULL;
  if (res->latent_state_chain) {
    cur_se = (double *)malloc(sizeof(double) * se_len);
    prop_se = (double *)malloc(sizeof(double) * se_len);
  }
  st = run_filter_se(c, y, cur, res->target_n, ORC_SISAR, ORC_STRATIFIED, ((uint32_t)PH_MAIN << 28) | 0u, chain_id, &cur_ll, cur_se);
  if (st) { free(cur_se); free(prop_se); return st; }
  memcpy(res->theta_chain, cur, sizeof(double) * (size_t)p);
  res->loglike_chain[0] = cur_ll;
  if (cur_se) memcpy(res->latent_state_chain, cur_se, sizeof(double) * se_len);
  res->n_accept = 0;
  for (int it = 1; it < c->m; it++) {
    double xi[8];
    for (int j = 0; j < p; j++) {
      zc[j] = orc_transform(cur[j], c->transform[j]);
      xi[j] = orc_noise_normal(seed, (uint32_t)PH_MAIN << 28, chain_id, (uint32_t)it, TAG_THETA_Z, 0, (uint32_t)j);
    }
    for (int a = 0; a < p; a++) {
      double s = zc[a];
      for (int b = 0; b <= a; b++) s += res->proposal_chol[a * p + b] * xi[b];
      zp[a] = s;
      prop[a] = orc_back_transform(zp[a], c->transform[a]);
    }
    double lp_prop_sum, lp_cur_sum;
    if (!priors_finite(c, prop, &lp_prop_sum)) { /* :435-442 reject without running the filter */
      memcpy(res->theta_chain + (size_t)it * p, cur, sizeof(double) * (size_t)p);
      res->loglike_chain[it] = cur_ll;
      if (cur_se) memcpy(res->latent_state_chain + (size_t)it * se_len, cur_se, sizeof(double) * se_len);
      continue;
    }
    double prop_ll;
    st = run_filter_se(c, y, prop, res->target_n, ORC_SISAR, ORC_STRATIFIED, ((uint32_t)PH_MAIN << 28) | (uint32_t)it, chain_id, &prop_ll, prop_se);
    if (st) { free(cur_se); free(prop_se); return st; }
    priors_finite(c, cur, &lp_cur_sum);
    double num = prop_ll + lp_prop_sum + orc_log_jacobian(prop, c->transform, p); /* :474-483 */
    double den = cur_ll + lp_cur_sum + orc_log_jacobian(cur, c->transform, p);
    double ratio = num - den;
    if (isnan(ratio)) ratio = -INFINITY;
    double u = orc_noise_uniform(seed, (uint32_t)PH_MAIN << 28, chain_id, (uint32_t)it, TAG_THETA_U, 0, 0);
    if (log(u) < ratio) {
      memcpy(cur, prop, sizeof(double) * (size_t)p); cur_ll = prop_ll; res->n_accept++;
      if (cur_se) memcpy(cur_se, prop_se, sizeof(double) * se_len);
    }
    memcpy(res->theta_chain + (size_t)it * p, cur, sizeof(double) * (size_t)p);
    res->loglike_chain[it] = cur_ll;
    if (cur_se) memcpy(res->latent_state_chain + (size_t)it * se_len, cur_se, sizeof(double) * se_len);
  }
  free(cur_se); free(prop_se);
  return ORC_OK;
}

/* ------------------------------------------------------------------------- */
/* R's default RNG restated: Mersenne-Twister (Matsumoto & Nishimura 1998)   */
/* with R's seed scrambling (src/main/RNG.c) and inversion normals.          */
/* Base R is not part of the reference tree; used for data simulation and as */
/* the realistic-cost noise source of the CPU baseline only.                 */
/* ------------------------------------------------------------------------- */
#define MT_N 624
#define MT_M 397
void orc_rrng_set_seed(orc_rrng *r, uint32_t seed) {
  /* RNG_Init: 50 LCG scrambles, then fill i_seed[0..624]; FixupSeeds: dummy[0]=624 */
  for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
  for (int j = 0; j < MT_N + 1; j++) {
    seed = 69069u * seed + 1u;
    r->mt[j] = seed;
  }
  r->mt[0] = MT_N; /* dummy[0] = mti */
  r->mti = MT_N;
}
static uint32_t mt_genrand(orc_rrng *r) {
  uint32_t *mt = r->mt + 1; /* dummy+1 */
  uint32_t y;
  static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
  if (r->mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
    }
    for (; kk < MT_N - 1; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
    r->mti = 0;
  }
  y = mt[r->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
double orc_rrng_unif(orc_rrng *r) {
  double v = (double)mt_genrand(r) * 2.3283064365386963e-10; /* [0,1) */
  if (v <= 0.0) return 0.5 * 2.328306437080797e-10; /* fixup: open interval */
  if (1.0 - v <= 0.0) return 1.0 - 0.5 * 2.328306437080797e-10;
  return v;
}
/* qnorm: Wichura (1988) AS241 PPND16, as used by R's qnorm5 */
static double qnorm_as241(double p) {
  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                   45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                 133.14166789178437745) * r + 3.387132872796366608) /
          (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
               21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
             42.313330701600911252) * r + 1.0);
    return val;
  }
  r = (q < 0) ? p : 1.0 - p;
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r +
               1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
             4.6303378461565452959) * r + 1.42343711074968357734) /
          (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r +
               0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r +
             2.05319162663775882187) * r + 1.0);
  } else {
    r -= 5.0;
    val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r +
               0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r +
             5.4637849111641143699) * r + 6.6579046435011037772) /
          (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
               7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r +
             0.59983224388131900458) * r + 1.0);
  }
  return (q < 0.0) ? -val : val;
}
double orc_rrng_norm(orc_rrng *r) { /* norm_rand(), INVERSION */
  const double BIG = 134217728.0;
  double u = orc_rrng_unif(r);
  u = (double)(int)(BIG * u) + orc_rrng_unif(r);
  return qnorm_as241(u / BIG);
}

/* ------------------------------------------------------------------------- */
/* CPU baseline: bootstrap filter driven by the R-like RNG above, no noise   */
/* injection, single thread (the R interpreter is single-threaded).          */
/* ------------------------------------------------------------------------- */
double orc_bench_bootstrap_filter(int model, int N, int T, const double *y, const double *theta,
                                  int resample_algorithm, int resample_fn, double threshold,
                                  uint32_t seed, double *loglike_out, int *n_resampled_out) {
  model_dims md;
  if (get_dims(model, &md) || md.d != 1 || md.nu_trans) return -1.0;
  orc_rrng rng;
  orc_rrng_set_seed(&rng, seed);
  double *px = (double *)malloc(sizeof(double) * (size_t)N);
  double *lw = (double *)malloc(sizeof(double) * (size_t)N);
  double *w = (double *)malloc(sizeof(double) * (size_t)N);
  double *tmp = (double *)malloc(sizeof(double) * (size_t)N);
  double *u = (double *)malloc(sizeof(double) * (size_t)N);
  int32_t *anc = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
  if (threshold < 0) threshold = resample_algorithm == ORC_SIS ? INFINITY : (resample_algorithm == ORC_SISR ? N : N / 2.0);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  double z;
  for (int i = 0; i < N; i++) { z = orc_rrng_norm(&rng); model_init(model, &px[i], theta, &z, NULL); }
  double loglike = 0.0; int nres = 0;
  for (int t = 0; t < T; t++) {
    for (int i = 0; i < N; i++) { z = orc_rrng_norm(&rng); model_transition(model, &px[i], theta, t + 1, &z, NULL, NULL); }
    for (int i = 0; i < N; i++) lw[i] = model_loglik(model, &y[t], &px[i], theta, t + 1);
    double mx = -INFINITY;
    for (int i = 0; i < N; i++) if (lw[i] > mx) mx = lw[i];
    if (mx < -1e8) { loglike = -INFINITY; break; }
    for (int i = 0; i < N; i++) tmp[i] = exp(lw[i] - mx);
    double s = r_sum(tmp, N);
    for (int i = 0; i < N; i++) w[i] = tmp[i] / s;
    loglike += mx + log(s) - log((double)N);
    for (int i = 0; i < N; i++) tmp[i] = w[i] * w[i];
    double ess = 1.0 / r_sum(tmp, N);
    int should = resample_algorithm == ORC_SIS ? 0 : (resample_algorithm == ORC_SISR ? 1 : ess < threshold);
    if (should) {
      if (resample_fn == ORC_SYSTEMATIC) orc_resample_systematic(N, w, orc_rrng_unif(&rng), anc);
      else {
        for (int i = 0; i < N; i++) u[i] = orc_rrng_unif(&rng);
        if (resample_fn == ORC_STRATIFIED) orc_resample_stratified(N, w, u, anc);
        else orc_resample_multinomial_rcpp(N, w, u, anc);
      }
      for (int i = 0; i < N; i++) tmp[i] = px[anc[i] - 1];
      memcpy(px, tmp, sizeof(double) * (size_t)N);
      for (int i = 0; i < N; i++) w[i] = 1.0 / N;
      nres++;
    }
    for (int i = 0; i < N; i++) tmp[i] = px[i] * w[i];
    volatile double se = r_sum(tmp, N);
    (void)se;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (loglike_out) *loglike_out = loglike;
  if (n_resampled_out) *n_resampled_out = nres;
  free(px); free(lw); free(w); free(tmp); free(u); free(anc);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
