// bssm_diag.cuh -- MCMC diagnostics of the PMMH draws on the device: multi-chain effective sample size
// (R/ESS.R:30-104: between/within variances, stats::acf per chain, Geyer initial monotone sequence) and
// split R-hat (R/rhat.R:27-67, with its [0.99, 1] -> 1 clamp), for all parameters of all chains at once
// (the reference calls both per parameter on every pmmh() return, R/pmmh.R:570-594).
//
// Five small kernels, none of which needs a barrier or a shuffle: every thread owns one output and walks
// its inputs in index order, so the lagged sums are accumulated exactly as stats::acf accumulates them
// (one double running sum per lag, src/library/stats/src/filter.c; R itself is not under /root/reference).
//   moments  thread (parameter, chain, role)  mean and variance of the whole chain / first half / second half;
//                                             the whole-chain thread also writes the centred series
//   acov     thread (parameter, chain, lag)   sum_i xc[i] xc[i+lag] / m; a warp covers 32 consecutive lags, so
//                                             xc[i] is a broadcast and xc[i+lag] one coalesced line
//   between  thread (parameter)               b, w, var_hat for ess(); the whole of split R-hat
//   rho      thread (parameter, lag)          hat_rho[lag] = 1 - (w - mean_c var_c acf_c[lag]) / var_hat
//   geyer    thread (parameter)               pairs, running minimum, sum to the first negative pair, ess
// The per-thread bodies are host+device: nvcc wraps them in __global__ launchers below, g++ runs the same text
// thread by thread in tests/host_diag.cpp (logic test without a GPU).
#pragma once
#ifndef BSSM_HD
#ifdef __CUDACC__
#define BSSM_HD __host__ __device__ __forceinline__
#else
#define BSSM_HD inline
#endif
#endif
#include <math.h>

namespace bssm {

struct DiagArgs {
  // draw of (chain c, kept iteration i, parameter j) = x[c * chain_stride + i * iter_stride + j]
  const double* x;
  long long chain_stride, iter_stride;
  int k, m, p;     // chains, kept iterations, parameters
  double* xc;      // [p][k][m] centred series
  double* mom;     // [p][k][3][2] (mean, variance) of the whole chain, first half, second half
  double* acov;    // [p][k][m] biased autocovariance
  double* par;     // [p][2] w, var_hat of ess()
  double* rho;     // [p][m]
  double* ess;     // [p] or nullptr
  double* rhat;    // [p]
  int* flags;      // [p] bit 0: a chain has zero variance (ess NA); bit 1: a half chain has (rhat NA)
};

BSSM_HD long long diag_n_moments(const DiagArgs& a) { return (long long)a.p * a.k * 3; }
BSSM_HD long long diag_n_acov(const DiagArgs& a) { return (long long)a.p * a.k * a.m; }
BSSM_HD long long diag_n_rho(const DiagArgs& a) { return (long long)a.p * a.m; }

// colMeans / apply(mat, 2, var) of R/ESS.R:45,52 and R/rhat.R:48,52 (two passes, as R's cov does), and the
// x - mean(x) that stats::acf applies before its lagged sums
BSSM_HD void diag_moments_thread(const DiagArgs& a, long long tid) {
  const int role = (int)(tid % 3);
  const long long s = tid / 3;              // series j * k + c
  const int c = (int)(s % a.k), j = (int)(s / a.k);
  const int mh = a.m / 2;                   // R/rhat.R:35-38: an odd last iteration is dropped
  const int lo = role == 2 ? mh : 0;
  const int n = role == 0 ? a.m : mh;
  const double* src = a.x + (long long)c * a.chain_stride + j;
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += src[(long long)(lo + i) * a.iter_stride];
  const double mean = sum / (double)n;
  double ss = 0.0;
  for (int i = 0; i < n; i++) {
    const double dlt = src[(long long)(lo + i) * a.iter_stride] - mean;
    ss += dlt * dlt;
    if (role == 0) a.xc[s * a.m + i] = dlt;
  }
  a.mom[(s * 3 + role) * 2 + 0] = mean;
  a.mom[(s * 3 + role) * 2 + 1] = ss / (double)(n - 1);
}

// acf(mat[, i], lag.max = m - 1)$acf numerator (R/ESS.R:66): sum over i of x[i] x[i + lag], divided by m
BSSM_HD void diag_acov_thread(const DiagArgs& a, long long tid) {
  const int lag = (int)(tid % a.m);
  const long long s = tid / a.m;
  const double* xc = a.xc + s * a.m;
  const int n = a.m - lag;
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += xc[i + lag] * xc[i];
  a.acov[s * a.m + lag] = sum / (double)a.m;
}

// R/ESS.R:45-61 (b, w, var_hat) and the whole of R/rhat.R:33-66
BSSM_HD void diag_between_thread(const DiagArgs& a, long long j) {
  const int k = a.k, m = a.m;
  int flag = 0;
  {
    double msum = 0.0, vsum = 0.0;
    for (int c = 0; c < k; c++) {
      const double* mo = a.mom + ((j * k + c) * 3 + 0) * 2;
      msum += mo[0];
      vsum += mo[1];
      if (mo[1] == 0.0) flag |= 1;
    }
    const double overall = msum / (double)k;
    double bs = 0.0;
    for (int c = 0; c < k; c++) {
      const double dlt = a.mom[((j * k + c) * 3 + 0) * 2] - overall;
      bs += dlt * dlt;
    }
    const double b = (double)m / (double)(k - 1) * bs;
    const double w = vsum / (double)k;
    a.par[j * 2 + 0] = w;
    a.par[j * 2 + 1] = ((double)(m - 1) / (double)m) * w + (1.0 / (double)m) * b;
  }
  {
    const int m2 = (m / 2) * 2, k2 = 2 * k;
    double msum = 0.0, vsum = 0.0;
    for (int h = 0; h < k2; h++) {           // chains_split column order: (chain, half), R/rhat.R:41-46
      const double* mo = a.mom + ((j * k + h / 2) * 3 + 1 + (h & 1)) * 2;
      msum += mo[0];
      vsum += mo[1];
      if (mo[1] == 0.0) flag |= 2;
    }
    const double overall = msum / (double)k2;
    double bs = 0.0;
    for (int h = 0; h < k2; h++) {
      const double dlt = a.mom[((j * k + h / 2) * 3 + 1 + (h & 1)) * 2] - overall;
      bs += dlt * dlt;
    }
    const double b = (double)m2 / (double)(k2 - 1) * bs;     // R/rhat.R:51: the full (even) length, not the half
    const double w = vsum / (double)k2;
    const double var_hat = ((double)(m2 - 1) / (double)m2) * w + (1.0 / (double)m2) * b;
    double r = sqrt(var_hat / w);
    if (r >= 0.99 && r <= 1.0) r = 1.0;                      // R/rhat.R:63-65
    a.rhat[j] = (flag & 2) ? nan("") : r;
  }
  a.flags[j] = flag;
}

// R/ESS.R:70-74.  acf = acov[lag] / (se * se), se = sqrt(acov[0]), kept inside [-1, 1] as stats::acf does
BSSM_HD void diag_rho_thread(const DiagArgs& a, long long tid) {
  const int lag = (int)(tid % a.m);
  const long long j = tid / a.m;
  double term = 0.0;
  for (int c = 0; c < a.k; c++) {
    const long long s = j * a.k + c;
    const double se = sqrt(a.acov[s * a.m]);
    double r = a.acov[s * a.m + lag] / (se * se);
    r = r > 1.0 ? 1.0 : (r < -1.0 ? -1.0 : r);
    term += a.mom[(s * 3 + 0) * 2 + 1] * r;
  }
  term *= 1.0 / (double)a.k;
  a.rho[j * a.m + lag] = 1.0 - (a.par[j * 2 + 0] - term) / a.par[j * 2 + 1];
}

// R/ESS.R:76-103: pairs of consecutive lags, made non-increasing, summed up to the first negative one
BSSM_HD void diag_geyer_thread(const DiagArgs& a, long long j) {
  if (!a.ess) return;
  const double* rho = a.rho + j * a.m;
  const int max_pairs = (a.m - 1) / 2;
  double sum_rho = 0.0, prev = 0.0;
  for (int t = 1; t <= max_pairs; t++) {
    double pr = rho[2 * t - 1] + rho[2 * t];
    if (t >= 2 && pr > prev) pr = prev;
    if (pr < 0.0) break;
    sum_rho += pr;
    prev = pr;
  }
  const double tau = 1.0 + 2.0 * sum_rho;
  a.ess[j] = (a.flags[j] & 1) ? nan("") : ((double)a.k * (double)a.m) / tau;
}

#ifdef __CUDACC__
#define BSSM_DIAG_KERNEL(name, body, count)                                                  \
  __global__ void __launch_bounds__(256) name(DiagArgs a) {                                  \
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;                  \
    if (tid < (count)) body(a, tid);                                                         \
  }
BSSM_DIAG_KERNEL(k_diag_moments, diag_moments_thread, diag_n_moments(a))
BSSM_DIAG_KERNEL(k_diag_acov, diag_acov_thread, diag_n_acov(a))
BSSM_DIAG_KERNEL(k_diag_between, diag_between_thread, (long long)a.p)
BSSM_DIAG_KERNEL(k_diag_rho, diag_rho_thread, diag_n_rho(a))
BSSM_DIAG_KERNEL(k_diag_geyer, diag_geyer_thread, (long long)a.p)
#undef BSSM_DIAG_KERNEL
#endif

}  // namespace bssm
