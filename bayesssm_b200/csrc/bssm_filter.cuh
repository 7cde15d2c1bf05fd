// bssm_filter.cuh -- general particle-filter kernels, batched over
// C independent filters (grid.x = filter, grid.y = block within the filter).  Replaces the per-observation loop of
// .particle_filter_core (R/particle_filter_core.R:76-246) for BPF / APF / RMPF, every built-in
// or NVRTC model, both precisions, injected or Philox noise, optional histories.
// The throughput configurations of the bootstrap filter use the persistent kernel in
// bssm_fast.cuh instead; this path is the general one and the parity reference.
#pragma once
#include "bssm_common.cuh"
#include "bssm_models.cuh"

namespace bssm {

constexpr int FT_THREADS = 256;
constexpr int PART_W = 8;  // doubles per block partial: m, s, q, sx[0..3], nan

// injected noise (device copies, double, particle index fastest); nullptr members => Philox
struct NoiseDev {
  const double *z_init, *u_init, *z_trans, *u_trans, *z_trans2, *u_trans2;
  const double *u_resample, *u_resample_aux, *z_move, *u_move;
  int injected;
};

// per-batch device arrays (SoA over filters)
struct FilterDev {
  int C, N, T, dy, d;          // N = stride / max particles
  const int* n_per;            // [C] particles per filter or nullptr (= N)
  const double* theta;         // [C][theta_stride]
  int theta_stride;
  const double* y;             // [T][dy]
  const int* obs_times;        // [T] or nullptr
  const unsigned int* stream;  // [C] Philox stream ids
  const unsigned int* run_id;  // [C]
  unsigned long long seed;
  NoiseDev noise;
  // state
  void *xa, *xb;               // [C][d][N] Real
  void *lw, *lw_aux, *auxg;    // [C][N] Real
  double* part;                // [C][nblk][PART_W]
  int nblk;
  double *M, *S, *loglike, *cur_ess;  // [C]
  int *alive, *resample, *status, *early_exit, *n_resampled, *cur;  // [C]
  // outputs
  double *ess;                 // [C][T+1]
  double *state_est;           // [C][T+1][d]
  double *loglike_history;     // [C][T] or nullptr
  double *particles_history;   // [C][hist_rows][d][N] or nullptr
  double *weights_history;     // [C][hist_rows][N] or nullptr
  int hist_rows;               // rows the device history buffers hold: a ring of 2 (row r lives in slot r % 2 until it has been copied out)
  int *anc_history;            // [C][T][N] 1-based (tests) or nullptr
  int *anc_aux_history;        // [C][T][N] or nullptr
  int algorithm, ralg;
  double threshold;            // < 0: reference default
  int carry;                   // 1: carried-weights mode (a stated DEVIATION from the reference, which drops the weights on steps that do
                               // not resample -- SURVEY App. A1): w_t ~ w_{t-1} g_t there; general kernels, BPF / RMPF
};

__device__ __forceinline__ int filt_n(const FilterDev& f, int c) { return f.n_per ? f.n_per[c] : f.N; }

// ---- noise access ---------------------------------------------------------------------------
template <typename Real>
__device__ __forceinline__ Real noise_normal(const FilterDev& f, const NoiseKey& key, const double* buf, int nslot,
                                             unsigned int tag, unsigned int t, int row, int slot, int i) {
  if (f.noise.injected) return (Real)buf[((size_t)row * nslot + slot) * f.N + i];
  uint4x q = noise_quad(key, t, tag, (unsigned int)slot, (unsigned int)i >> 2);
  int pr = (i & 3) >> 1;
  Real n0, n1;
  Math<Real>::box_muller(q.w[2 * pr], q.w[2 * pr + 1], n0, n1);
  return (i & 1) ? n1 : n0;
}
__device__ __forceinline__ double noise_uniform(const FilterDev& f, const NoiseKey& key, const double* buf, int nslot,
                                                unsigned int tag, unsigned int t, int row, int slot, int i) {
  if (f.noise.injected) return buf[((size_t)row * nslot + slot) * f.N + i];
  uint4x q = noise_quad(key, t, tag, (unsigned int)slot, (unsigned int)i >> 2);
  return word_to_unit_f64(q.w[i & 3]);
}

// ---- online (max, sum exp, sum exp^2, sum exp*x) accumulator -------------------------------------
struct Acc {
  double m, s, q, sx[4];
  int nan;
};
__device__ __forceinline__ void acc_init(Acc& a) {
  a.m = -__longlong_as_double(0x7FF0000000000000LL); a.s = 0; a.q = 0; a.nan = 0;
  a.sx[0] = a.sx[1] = a.sx[2] = a.sx[3] = 0;
}
// F32: the rescaling factors by the SFU exp of the throughput precision (the per-particle terms were formed that way too); the
// fp64 exp costs ~5x the rest of a merge, and a block reduction merges ten times per thread
template <bool F32 = false>
__device__ __forceinline__ void acc_merge(Acc& a, const Acc& b, int d) {
  a.nan |= b.nan;
  if (b.s == 0.0 && b.q == 0.0 && !(b.m > a.m) && b.sx[0] == 0.0 && b.sx[1] == 0.0 && b.sx[2] == 0.0 && b.sx[3] == 0.0) return;
  double m = a.m > b.m ? a.m : b.m;
  if (m == -__longlong_as_double(0x7FF0000000000000LL)) { a.m = m; return; }
  double ea, eb;
  if constexpr (F32) { ea = (a.m == m) ? 1.0 : (double)Math<float>::exp_((float)(a.m - m)); eb = (b.m == m) ? 1.0 : (double)Math<float>::exp_((float)(b.m - m)); }
  else { ea = (a.m == m) ? 1.0 : exp(a.m - m); eb = (b.m == m) ? 1.0 : exp(b.m - m); }
  a.s = a.s * ea + b.s * eb;
  a.q = a.q * (ea * ea) + b.q * (eb * eb);
  for (int k = 0; k < d; k++) a.sx[k] = a.sx[k] * ea + b.sx[k] * eb;
  a.m = m;
}
template <typename Real>
__device__ __forceinline__ void acc_add(Acc& a, Real lw, const Real* x, int d) {
  if (lw != lw) { a.nan = 1; return; }
  double l = (double)lw;
  if (l == -__longlong_as_double(0x7FF0000000000000LL)) return;  // exp(-inf - m) = 0
  if (l <= a.m) {
    double e = (double)Math<Real>::exp_((Real)(l - a.m));
    a.s += e; a.q += e * e;
    for (int k = 0; k < d; k++) a.sx[k] += e * (double)x[k];
  } else {
    double r = (a.s == 0.0 && a.q == 0.0) ? 0.0 : (double)Math<Real>::exp_((Real)(a.m - l));
    a.s = a.s * r + 1.0; a.q = a.q * (r * r) + 1.0;
    for (int k = 0; k < d; k++) a.sx[k] = a.sx[k] * r + (double)x[k];
    a.m = l;
  }
}
__device__ __forceinline__ Acc acc_shfl_down(const Acc& a, int o) {
  Acc b;
  b.m = __shfl_down_sync(0xffffffffu, a.m, o); b.s = __shfl_down_sync(0xffffffffu, a.s, o);
  b.q = __shfl_down_sync(0xffffffffu, a.q, o);
  for (int k = 0; k < 4; k++) b.sx[k] = __shfl_down_sync(0xffffffffu, a.sx[k], o);
  b.nan = __shfl_down_sync(0xffffffffu, a.nan, o);
  return b;
}
// deterministic block reduction (fixed tree); result valid in thread 0
template <bool F32 = false>
__device__ __forceinline__ void acc_block_reduce(Acc& a, int d, Acc* sm /* >= 32 */) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int o = 16; o; o >>= 1) { Acc b = acc_shfl_down(a, o); if (lane + o < 32) acc_merge<F32>(a, b, d); }
  __syncthreads();
  if (lane == 0) sm[wid] = a;
  __syncthreads();
  if (wid == 0) {
    Acc t;
    if (lane < nw) t = sm[lane]; else acc_init(t);
    for (int o = 16; o; o >>= 1) { Acc b = acc_shfl_down(t, o); if (lane + o < 32) acc_merge<F32>(t, b, d); }
    a = t;
  }
}
__device__ __forceinline__ void acc_store(const Acc& a, double* p) {
  p[0] = a.m; p[1] = a.s; p[2] = a.q; p[3] = a.sx[0]; p[4] = a.sx[1]; p[5] = a.sx[2]; p[6] = a.sx[3]; p[7] = (double)a.nan;
}
__device__ __forceinline__ void acc_load(Acc& a, const double* p) {
  a.m = p[0]; a.s = p[1]; a.q = p[2]; a.sx[0] = p[3]; a.sx[1] = p[4]; a.sx[2] = p[5]; a.sx[3] = p[6]; a.nan = (int)p[7];
}

template <typename Real> __device__ __forceinline__ Real* x_cur(const FilterDev& f, int c) {
  Real* base = (Real*)(f.cur[c] ? f.xb : f.xa);
  return base + (size_t)c * f.d * f.N;
}
template <typename Real> __device__ __forceinline__ Real* x_other(const FilterDev& f, int c) {
  Real* base = (Real*)(f.cur[c] ? f.xa : f.xb);
  return base + (size_t)c * f.d * f.N;
}

// ---- K1: init (R/particle_filter_core.R:76-116) ------------------------------------------------
template <typename Model, typename Real>
__global__ void __launch_bounds__(FT_THREADS) k_init(FilterDev f) {
  __shared__ Acc sm[32];
  int c = blockIdx.x;   // filter in grid.x (no 65535 limit), block within the filter in grid.y
  if (!f.alive[c]) return;
  int n = filt_n(f, c);
  Real par[Model::NPAR];
  Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
  NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  Real* x = (Real*)f.xa + (size_t)c * f.d * f.N;
  Acc a; acc_init(a); a.m = 0.0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    Real z[Model::NZ_INIT > 0 ? Model::NZ_INIT : 1];
    double u[Model::NU_INIT > 0 ? Model::NU_INIT : 1];
    for (int s = 0; s < Model::NZ_INIT; s++) z[s] = noise_normal<Real>(f, key, f.noise.z_init, Model::NZ_INIT, TAG_INIT_Z, T_INIT, 0, s, i);
    for (int s = 0; s < Model::NU_INIT; s++) u[s] = noise_uniform(f, key, f.noise.u_init, Model::NU_INIT, TAG_INIT_U, T_INIT, 0, s, i);
    Real xi[Model::D];
    Model::template init<Real>(xi, par, z, u);
    for (int k = 0; k < Model::D; k++) { x[(size_t)k * f.N + i] = xi[k]; a.sx[k] += (double)xi[k]; }
    a.s += 1.0;
  }
  acc_block_reduce<sizeof(Real) == 4>(a, Model::D, sm);
  if (threadIdx.x == 0) acc_store(a, f.part + ((size_t)c * f.nblk + blockIdx.y) * PART_W);
}

// ---- K2: propagate + log-weight + block partials ---------------------------------------------
// flags: GAP = run the gap transitions of observation `obs` (R/particle_filter_core.R:125-136)
//        SECOND = APF second transition at the same t (:159)
// wkind: 0 = log_likelihood (:177-183), 1 = aux log-likelihood into lw_aux (:142-147),
//        2 = log_likelihood - gathered aux (:169-175)
enum { WF_GAP = 1, WF_SECOND = 2 };
template <typename Model, typename Real>
__global__ void __launch_bounds__(FT_THREADS, 3) k_weight(FilterDev f, int obs, int flags, int wkind) {
  __shared__ Acc sm[32];
  int c = blockIdx.x;   // filter in grid.x (no 65535 limit), block within the filter in grid.y
  if (!f.alive[c]) return;
  int n = filt_n(f, c);
  Real par[Model::NPAR];
  Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
  NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  Real* x = x_cur<Real>(f, c);
  Real* lw = (Real*)(wkind == 1 ? f.lw_aux : f.lw) + (size_t)c * f.N;
  const Real* auxg = (const Real*)f.auxg + (size_t)c * f.N;
  int ot = f.obs_times ? f.obs_times[obs] : obs + 1;
  int prev_t = obs == 0 ? 0 : (f.obs_times ? f.obs_times[obs - 1] : obs);
  double yv[4];
  for (int k = 0; k < f.dy && k < 4; k++) yv[k] = f.y[(size_t)obs * f.dy + k];
  Acc a; acc_init(a);
  const bool carried = f.carry && wkind == 0 && obs > 0 && !f.resample[c];   // (f.resample[c]: still the previous observation's decision)
  const double m_prev = carried ? f.M[c] : 0.0, s_prev = carried ? f.S[c] : 1.0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    Real xi[Model::D];
    for (int k = 0; k < Model::D; k++) xi[k] = x[(size_t)k * f.N + i];
    Real z[Model::NZ_TRANS > 0 ? Model::NZ_TRANS : 1];
    double u[Model::NU_TRANS > 0 ? Model::NU_TRANS : 1];
    if (flags & WF_GAP) {
      for (int tnow = prev_t + 1; tnow <= ot; tnow++) {
        if constexpr (ModelDynU<Model>::value) {      // uniforms on demand (Philox mode only: checked on the host)
          DynU du(key, (unsigned int)(tnow - 1), TAG_TRANS_DYN, (unsigned int)i);
          Model::template transition_dyn<Real>(xi, par, tnow, du);
          continue;
        }
        for (int s = 0; s < Model::NZ_TRANS; s++) z[s] = noise_normal<Real>(f, key, f.noise.z_trans, Model::NZ_TRANS, TAG_TRANS_Z, (unsigned int)(tnow - 1), tnow - 1, s, i);
        for (int s = 0; s < Model::NU_TRANS; s++) u[s] = noise_uniform(f, key, f.noise.u_trans, Model::NU_TRANS, TAG_TRANS_U, (unsigned int)(tnow - 1), tnow - 1, s, i);
        Model::template transition<Real>(xi, par, tnow, z, u);
      }
    }
    if (flags & WF_SECOND) {
      if constexpr (ModelDynU<Model>::value) {
        DynU du(key, (unsigned int)obs, TAG_TRANS2_DYN, (unsigned int)i);
        Model::template transition_dyn<Real>(xi, par, ot, du);
      } else {
        for (int s = 0; s < Model::NZ_TRANS; s++) z[s] = noise_normal<Real>(f, key, f.noise.z_trans2, Model::NZ_TRANS, TAG_TRANS2_Z, (unsigned int)obs, obs, s, i);
        for (int s = 0; s < Model::NU_TRANS; s++) u[s] = noise_uniform(f, key, f.noise.u_trans2, Model::NU_TRANS, TAG_TRANS2_U, (unsigned int)obs, obs, s, i);
        Model::template transition<Real>(xi, par, ot, z, u);
      }
    }
    if (flags) for (int k = 0; k < Model::D; k++) x[(size_t)k * f.N + i] = xi[k];
    Real l;
    if (wkind == 1) l = Model::template aux_loglik<Real>(yv, xi, par, ot);
    else {
      l = Model::template loglik<Real>(yv, xi, par, ot);
      if (wkind == 2) l = l - auxg[i];
      // carried weights: + log(n W_{t-1,i}), W the previous observation's normalised weight, when it did not resample (after a
      // resampling, and at the first observation, every W is 1 / n)
      if (carried) l = l + (Real)log((double)n * (exp((double)lw[i] - m_prev) / s_prev));
    }
    lw[i] = l;
    acc_add<Real>(a, l, xi, Model::D);
  }
  acc_block_reduce<sizeof(Real) == 4>(a, Model::D, sm);
  if (threadIdx.x == 0) acc_store(a, f.part + ((size_t)c * f.nblk + blockIdx.y) * PART_W);
}

// ---- K3: per-filter finalise -------------------------------------------------------------------
// kind 0: t = 0 bookkeeping (:106-116); 1: main weights (:189-224); 2: APF first stage (:152-155);
// 3: after resampling (:222-223,237-241)
// F32: the throughput precision merges the block records with the SFU exp (acc_merge)
template <bool F32>
static __global__ void __launch_bounds__(128) k_finalize(FilterDev f, int obs, int kind) {
  __shared__ Acc sm[32];
  int c = blockIdx.x;
  if (!f.alive[c]) return;
  if (kind == 3 && !f.resample[c]) return;
  int n = filt_n(f, c);
  int nb = f.nblk;
  Acc a; acc_init(a);
  if (kind == 0 || kind == 3) a.m = 0.0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    Acc t; acc_load(t, f.part + ((size_t)c * f.nblk + b) * PART_W);
    acc_merge<F32>(a, t, f.d);
  }
  acc_block_reduce<F32>(a, f.d, sm);
  if (threadIdx.x) return;
  const int T1 = f.T + 1;
  if (kind == 0) {
    f.ess[(size_t)c * T1] = (double)n;
    for (int k = 0; k < f.d; k++) f.state_est[((size_t)c * T1) * f.d + k] = a.sx[k] / (double)n;
    return;
  }
  if (kind == 3) {
    f.ess[(size_t)c * T1 + obs + 1] = (double)n;
    for (int k = 0; k < f.d; k++) f.state_est[((size_t)c * T1 + obs + 1) * f.d + k] = a.sx[k] / (double)n;
    f.n_resampled[c] += 1;
    return;
  }
  if (a.nan) { f.status[c] = 3 /*BSSM_ERR_NAN_WEIGHT*/; f.alive[c] = 0; f.resample[c] = 0; return; }
  f.M[c] = a.m; f.S[c] = a.s;
  if (kind == 2) { f.resample[c] = 1; return; }
  if (a.m < -1e8) {  // all(lw < -1e8): R/particle_filter_core.R:189-202
    double ninf = -__longlong_as_double(0x7FF0000000000000LL);
    f.loglike[c] = ninf;
    if (f.loglike_history) f.loglike_history[(size_t)c * f.T + obs] = ninf;
    f.early_exit[c] = 1; f.alive[c] = 0; f.resample[c] = 0;
    return;
  }
  double ll = f.loglike[c] + (a.m + log(a.s) - log((double)n));
  f.loglike[c] = ll;
  if (f.loglike_history) f.loglike_history[(size_t)c * f.T + obs] = ll;
  double ess = (a.s * a.s) / a.q;
  f.ess[(size_t)c * T1 + obs + 1] = ess;
  double thr = f.threshold;
  int ralg = (f.algorithm == 2) ? 1 : f.ralg;
  if (thr < 0 || f.algorithm == 2) thr = (ralg == 0) ? __longlong_as_double(0x7FF0000000000000LL) : (ralg == 1 ? (double)n : (double)n / 2.0);
  int should = (ralg == 0) ? 0 : (ralg == 1 ? 1 : (ess < thr));
  if (f.algorithm == 2) should = 1;
  f.resample[c] = should;
  if (!should)
    for (int k = 0; k < f.d; k++) f.state_est[((size_t)c * T1 + obs + 1) * f.d + k] = a.sx[k] / a.s;
}

// ---- K5+K7: index search + ancestor gather ----------------------------------------------------
struct USrcFilter {  // resampling uniforms of observation `obs`: injected [T][N] or Philox
  const double* buf; int N; int injected; unsigned long long seed; const unsigned int* run_id;
  const unsigned int* stream; unsigned int tag; int obs;
  __device__ __forceinline__ double operator()(int seg, int i) const {
    if (injected) return buf[(size_t)obs * N + i];
    NoiseKey key = make_key(seed, run_id[seg], stream[seg]);
    uint4x q = noise_quad(key, (unsigned int)obs, tag, 0u, (unsigned int)i >> 2);
    return word_to_unit_f64(q.w[i & 3]);
  }
};

template <typename Real>
__global__ void __launch_bounds__(FT_THREADS) k_search_gather(FilterDev f, USrcFilter us, int fn, int obs, int aux_stage,
                                                             const double* __restrict__ cdf) {
  int c = blockIdx.x;   // filter in grid.x (no 65535 limit), block within the filter in grid.y
  if (!f.alive[c] || !f.resample[c]) return;
  int n = filt_n(f, c);
  const double* cd = cdf + (size_t)c * f.N;
  const Real* xs = x_cur<Real>(f, c);
  Real* xd = x_other<Real>(f, c);
  const Real* lwa = (const Real*)f.lw_aux + (size_t)c * f.N;
  Real* ag = (Real*)f.auxg + (size_t)c * f.N;
  int* hist = aux_stage ? f.anc_aux_history : f.anc_history;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    double pos;
    if (fn == 1) pos = ((double)i + us(c, 0)) / (double)n;
    else { double u = us(c, i); pos = (fn == 0) ? ((double)i + u) / (double)n : u; }
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = lo + ((hi - lo) >> 1); if (cd[mid] < pos) lo = mid + 1; else hi = mid; }
    for (int k = 0; k < f.d; k++) xd[(size_t)k * f.N + i] = xs[(size_t)k * f.N + lo];
    if (aux_stage) ag[i] = lwa[lo];
    if (hist) hist[((size_t)c * f.T + obs) * f.N + i] = lo + 1;
  }
}
// flip the current-buffer flag after a gather
static __global__ void k_flip(FilterDev f) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < f.C && f.alive[c] && f.resample[c]) f.cur[c] ^= 1;
}

// ---- K8 + post-resample sums: RMPF move (R/particle_filter_core.R:226-234), sum x ---------------
template <typename Model, typename Real>
__global__ void __launch_bounds__(FT_THREADS) k_post(FilterDev f, int obs) {
  __shared__ Acc sm[32];
  int c = blockIdx.x;   // filter in grid.x (no 65535 limit), block within the filter in grid.y
  if (!f.alive[c] || !f.resample[c]) return;
  int n = filt_n(f, c);
  Real par[Model::NPAR];
  Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
  NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  Real* x = x_cur<Real>(f, c);
  int ot = f.obs_times ? f.obs_times[obs] : obs + 1;
  double yv[4];
  for (int k = 0; k < f.dy && k < 4; k++) yv[k] = f.y[(size_t)obs * f.dy + k];
  Acc a; acc_init(a); a.m = 0.0;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    Real xi[Model::D];
    for (int k = 0; k < Model::D; k++) xi[k] = x[(size_t)k * f.N + i];
    if (f.algorithm == 2 && Model::HAS_MOVE) {
      Real z[Model::NZ_MOVE > 0 ? Model::NZ_MOVE : 1];
      double u[Model::NU_MOVE > 0 ? Model::NU_MOVE : 1];
      for (int s = 0; s < Model::NZ_MOVE; s++) z[s] = noise_normal<Real>(f, key, f.noise.z_move, Model::NZ_MOVE, TAG_MOVE_Z, (unsigned int)obs, obs, s, i);
      for (int s = 0; s < Model::NU_MOVE; s++) u[s] = noise_uniform(f, key, f.noise.u_move, Model::NU_MOVE, TAG_MOVE_U, (unsigned int)obs, obs, s, i);
      Model::template move<Real>(xi, yv, par, ot, z, u);
      for (int k = 0; k < Model::D; k++) x[(size_t)k * f.N + i] = xi[k];
    }
    for (int k = 0; k < Model::D; k++) a.sx[k] += (double)xi[k];
  }
  acc_block_reduce<sizeof(Real) == 4>(a, Model::D, sm);
  if (threadIdx.x == 0) acc_store(a, f.part + ((size_t)c * f.nblk + blockIdx.y) * PART_W);
}

// ---- histories (R/particle_filter_core.R:100-116,242-245) ---------------------------------------
template <typename Real>
__global__ void __launch_bounds__(FT_THREADS) k_history(FilterDev f, int row /* 0..T */) {
  int c = blockIdx.x;   // filter in grid.x (no 65535 limit), block within the filter in grid.y
  int n = filt_n(f, c);
  const size_t slot = (size_t)c * f.hist_rows + row % f.hist_rows;
  if (!f.alive[c]) {
    // a filter that has stopped (early exit, R/particle_filter_core.R:189-202: the remaining rows stay zero): the ring slot is
    // copied out whatever happens, so it must not keep the row of two observations ago
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < f.N; i += gridDim.y * blockDim.x) {
      if (f.particles_history) for (int k = 0; k < f.d; k++) f.particles_history[(slot * f.d + k) * f.N + i] = 0.0;
      if (f.weights_history) f.weights_history[slot * f.N + i] = 0.0;
    }
    return;
  }
  const Real* x = x_cur<Real>(f, c);
  const Real* lw = (const Real*)f.lw + (size_t)c * f.N;
  bool uniform = (row == 0) || f.resample[c];
  double M = f.M[c], S = f.S[c];
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < f.N; i += gridDim.y * blockDim.x) {
    const bool on = i < n;   // ragged batches: the tail of a row beyond this filter's particles is zero
    if (f.particles_history)
      for (int k = 0; k < f.d; k++)
        f.particles_history[(slot * f.d + k) * f.N + i] = on ? (double)x[(size_t)k * f.N + i] : 0.0;
    if (f.weights_history)
      f.weights_history[slot * f.N + i] = !on ? 0.0 : (uniform ? 1.0 / (double)n : exp((double)lw[i] - M) / S);
  }
}

}  // namespace bssm
