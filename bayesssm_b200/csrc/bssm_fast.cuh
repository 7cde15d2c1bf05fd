// bssm_fast.cuh -- persistent bootstrap-filter kernel: the whole T loop of
// .particle_filter_core (R/particle_filter_core.R:123-246) for algorithm "BPF" in ONE launch.
//
// A filter is run by a GROUP of G co-resident CTAs (cooperative launch); CTA b owns the
// contiguous slice [b*nb, (b+1)*nb) of the particles and keeps it in REGISTERS for all T steps
// (PPT particles per thread).  Per observation:
//   P1  propagate (Philox normals, one Philox call per 4 particles) + log-weight, block max,
//       block sums of e = exp(lw - max_b), e^2, e*x                         [registers + shuffles]
//   B1  each CTA publishes (max_b, sums) as an epoch-stamped record in L2; every CTA polls the G
//       records (release/acquire, no atomics) and derives, redundantly but identically, the global
//       max / sum / ESS / resampling decision and its own cdf offset          [1 L2 round trip]
//   P3  local fp64 cdf of the normalised weights into shared memory (block scan)
//   P4  each CTA serves the contiguous run of OUTPUT slots whose positions (i + U_i)/N fall inside
//       its cdf interval: lower_bound in shared memory, coalesced store of the chosen x to x_new
//   B2  epoch-stamped "done" records; every CTA reloads its slice of x_new     [1 L2 round trip]
// Nothing but x_new (4 B write + 4 B read per particle, L2 resident) and the tiny records leaves
// the SM.  Same Philox keying and tie rule as the general engine, so results are independent of
// G and of the launch geometry up to floating-point summation order.
#pragma once
#include "bssm_common.cuh"
#include "bssm_filter.cuh"
#include "bssm_models.cuh"

namespace bssm {

constexpr int FAST_PPT = 8;          // particles per thread (registers)
constexpr int FAST_THREADS = 896;    // max threads per CTA (<= 72 registers each)
constexpr int FAST_MAX_NB = FAST_THREADS * FAST_PPT;  // 7168 particles per CTA
constexpr int FAST_MAX_G = 256;      // CTAs per group

struct __align__(64) FastRec {   // published once per observation by each CTA
  double m, s, q, sx;
  int nan; unsigned int epoch;
  double pad[3];
};
struct __align__(32) FastRec2 {  // published after the scatter of a resampling step
  double sumx;
  unsigned int epoch; int pad0;
  double pad[2];
};

struct FastParams {
  FilterDev f;
  int G, ngroups;
  int resample_fn;
  FastRec* rec;     // [ngroups][2][G]
  FastRec2* rec2;   // [ngroups][G]
  void* xnew;       // [ngroups][G * nb_max] Real
  int nb_max;       // slice stride (multiple of FAST_PPT)
};

__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename Real> struct FastMath;
template <> struct FastMath<float> {
  static __device__ __forceinline__ float exp_(float x) { return __expf(x); }
};
template <> struct FastMath<double> {
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// number of output slots i in [0, n) whose position is <= c, i.e. #{ i : pos_i <= c }.
// pos_i = (i + U_i) / n (stratified, U_i = Philox word of slot i) or (i + U) / n (systematic).
// Positions are non-decreasing in i, so this is the first i with pos_i > c.
struct PosGen {
  NoiseKey key; unsigned int obs; int fn; int n; bool pow2; double inv_n; double u_sys;
  __device__ __forceinline__ double u_of(int i) const {
    if (fn == 1) return u_sys;
    uint4x q = noise_quad(key, obs, TAG_RESAMP_U, 0u, (unsigned int)i >> 2);
    return word_to_unit_f64(q.w[i & 3]);
  }
  __device__ __forceinline__ double pos(int i, double u) const {
    double s = (double)i + u;
    return pow2 ? s * inv_n : s / (double)n;
  }
  __device__ __forceinline__ int count_le(double c) const {
    if (!(c > 0.0)) return 0;
    double t = c * (double)n;
    int i = (t >= (double)n) ? n - 1 : (int)t;   // candidate: stratum containing c
    // move down while slot i is above c, up while the next slot is still <= c
    while (i >= 0 && pos(i, u_of(i)) > c) i--;
    while (i + 1 < n && pos(i + 1, u_of(i + 1)) <= c) i++;
    return i + 1;
  }
};

// first j in [lo, n_loc-1] with cdf[j] >= p (clamped to n_loc-1); cdf in shared memory
__device__ __forceinline__ int smem_lower_bound(const double* cdf, int lo, int n_loc, double p) {
  int hi = n_loc - 1;
  if (lo >= hi || cdf[lo] >= p) return lo < hi ? lo : hi;
  lo++;
  if (lo >= hi || cdf[lo] >= p) return lo < hi ? lo : hi;
  lo++;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] < p) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <typename Model, typename Real>
__global__ void __launch_bounds__(FAST_THREADS, 1) k_fast_bpf(FastParams P) {
  static_assert(Model::D == 1 && Model::NZ_TRANS == 1 && Model::NU_TRANS == 0 && Model::NZ_INIT == 1 && Model::NU_INIT == 0,
                "persistent kernel: 1-D models with one normal per transition");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FilterDev& f = P.f;
  const int G = P.G;
  const int group = blockIdx.x / G, b = blockIdx.x % G;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
  // shared memory carve-up
  double* s_cdf = (double*)smem_raw;                               // [nb_max]
  double* s_tab = s_cdf + P.nb_max;                                // [4][G]: m, s, q, sx of every CTA
  double* s_red = s_tab + 4 * G;                                   // [4][32] reduction scratch
  Real* s_x = (Real*)(s_red + 4 * 32 + 16);                        // [nb_max]
  __shared__ int s_flag[4];
  __shared__ double s_misc[8];

  FastRec* rec = P.rec + (size_t)group * 2 * G;
  FastRec2* rec2 = P.rec2 + (size_t)group * G;
  Real* xnew = (Real*)P.xnew + (size_t)group * G * P.nb_max;
  unsigned int ep1 = 0, ep2 = 0;   // record epochs (B1 / B2): identical sequences in every CTA of the group
  if (tid < 4) s_flag[tid] = 0;
  __syncthreads();
  const double NINF = -__longlong_as_double(0x7FF0000000000000LL);

  for (int c = group; c < f.C; c += P.ngroups) {
    if (!f.alive[c]) continue;
    const int n = filt_n(f, c);
    int nb = (n + G - 1) / G;
    nb = (nb + FAST_PPT - 1) / FAST_PPT * FAST_PPT;
    const int base = b * nb;                                  // first global particle of this CTA
    const int n_loc = max(0, min(n - base, nb));              // particles owned by this CTA
    const int ibase = base + tid * FAST_PPT;                  // first global particle of this thread
    const bool pow2 = (n & (n - 1)) == 0;
    Real par[Model::NPAR];
    Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
    const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
    const int T1 = f.T + 1;

    // ---- init (R/particle_filter_core.R:76-116) ----
    Real x[FAST_PPT];
    double sum0 = 0.0;
#pragma unroll
    for (int h = 0; h < FAST_PPT / 4; h++) {
      uint4x qd = noise_quad(key, T_INIT, TAG_INIT_Z, 0u, (unsigned int)(ibase + 4 * h) >> 2);
      Real z0, z1, z2, z3;
      Math<Real>::box_muller(qd.w[0], qd.w[1], z0, z1);
      Math<Real>::box_muller(qd.w[2], qd.w[3], z2, z3);
      Real zz[4] = {z0, z1, z2, z3};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        Real xi[1]; Real zi[1] = {zz[k]};
        Model::template init<Real>(xi, par, zi, nullptr);
        x[4 * h + k] = xi[0];
        if (ibase + 4 * h + k < n && tid * FAST_PPT + 4 * h + k < nb) sum0 += (double)xi[0];
      }
    }
    // t = 0 outputs need the global mean: publish through the same record path
    double loglike = 0.0;
    int n_resampled = 0;
    bool dead = false;
    // block sum of sum0 -> record; CTA 0 gathers the t = 0 state estimate
    {
      double v = warp_sum_d(sum0);
      if (lane == 0) s_red[wid] = v;
      __syncthreads();
      if (wid == 0) {
        double t = lane < nw ? s_red[lane] : 0.0;
        t = warp_sum_d(t);
        if (lane == 0) {
          FastRec* r = &rec[((ep1 + 1) & 1) * G + b];
          r->m = 0.0; r->s = 0.0; r->q = 0.0; r->sx = t; r->nan = 0;
          __threadfence();
          st_release_u32(&r->epoch, ep1 + 1);
        }
      }
      ep1++;
      if (b == 0) {
        double v0 = 0.0;
        for (int j = tid; j < G; j += blockDim.x) {
          const FastRec* r = &rec[(ep1 & 1) * G + j];
          while (ld_acquire_u32(&r->epoch) != ep1) {}
          v0 += __ldcg(&r->sx);
        }
        v0 = warp_sum_d(v0);
        __syncthreads();
        if (lane == 0) s_red[wid] = v0;
        __syncthreads();
        if (tid == 0) {
          double t = 0.0;
          for (int w = 0; w < nw; w++) t += s_red[w];
          f.ess[(size_t)c * T1] = (double)n;
          f.state_est[(size_t)c * T1] = t / (double)n;
        }
      }
      __syncthreads();
    }

    for (int obs = 0; obs < f.T; obs++) {
      const int ot = f.obs_times ? f.obs_times[obs] : obs + 1;
      const int prev_t = obs == 0 ? 0 : (f.obs_times ? f.obs_times[obs - 1] : obs);
      double yv[4] = {0, 0, 0, 0};
      for (int k = 0; k < f.dy && k < 4; k++) yv[k] = f.y[(size_t)obs * f.dy + k];

      // ---- P1: propagate + log-weight ----
      Real lw[FAST_PPT];
      for (int tnow = prev_t + 1; tnow <= ot; tnow++) {
#pragma unroll
        for (int h = 0; h < FAST_PPT / 4; h++) {
          uint4x qd = noise_quad(key, (unsigned int)(tnow - 1), TAG_TRANS_Z, 0u, (unsigned int)(ibase + 4 * h) >> 2);
          Real z0, z1, z2, z3;
          Math<Real>::box_muller(qd.w[0], qd.w[1], z0, z1);
          Math<Real>::box_muller(qd.w[2], qd.w[3], z2, z3);
          Real zz[4] = {z0, z1, z2, z3};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            Real zi[1] = {zz[k]};
            Model::template transition<Real>(&x[4 * h + k], par, tnow, zi, nullptr);
          }
        }
      }
      Real mloc = Math<Real>::ninf();
      int nanf = 0;
#pragma unroll
      for (int k = 0; k < FAST_PPT; k++) {
        const bool own = (ibase + k < n) && (tid * FAST_PPT + k < nb);
        if (!own) x[k] = (Real)0;   // padding lanes never hold garbage (x_new beyond n is not written)
        Real l = Model::template loglik<Real>(yv, &x[k], par, ot);
        if (!own) l = Math<Real>::ninf();
        if (l != l) nanf = 1;
        lw[k] = l;
        mloc = l > mloc ? l : mloc;
      }
      // block max
      {
        Real v = mloc;
#pragma unroll
        for (int o = 16; o; o >>= 1) { Real t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
        nanf = __any_sync(0xffffffffu, nanf);
        if (lane == 0) { s_red[wid] = (double)v; s_red[32 + wid] = (double)nanf; }
        __syncthreads();
        if (wid == 0) {
          double t = lane < nw ? s_red[lane] : NINF;
          double nf = lane < nw ? s_red[32 + lane] : 0.0;
#pragma unroll
          for (int o = 16; o; o >>= 1) { double u = __shfl_xor_sync(0xffffffffu, t, o); t = u > t ? u : t; nf += __shfl_xor_sync(0xffffffffu, nf, o); }
          if (lane == 0) { s_misc[0] = t; s_misc[1] = nf; }
        }
        __syncthreads();
      }
      const double mb = s_misc[0];
      const int nan_b = s_misc[1] != 0.0;
      // e = exp(lw - mb); block sums
      Real e[FAST_PPT];
      double ts = 0.0, tq = 0.0, tx = 0.0;
      {
        Real fs = 0, fq = 0, fx = 0;
#pragma unroll
        for (int k = 0; k < FAST_PPT; k++) {
          Real ek = (lw[k] == Math<Real>::ninf() || mb == NINF) ? (Real)0 : FastMath<Real>::exp_(lw[k] - (Real)mb);
          e[k] = ek;
          fs += ek; fq += ek * ek; fx += ek * x[k];
        }
        ts = warp_sum_d((double)fs); tq = warp_sum_d((double)fq); tx = warp_sum_d((double)fx);
        if (lane == 0) { s_red[wid] = ts; s_red[32 + wid] = tq; s_red[64 + wid] = tx; }
        __syncthreads();
        if (wid == 0) {
          double a0 = lane < nw ? s_red[lane] : 0.0, a1 = lane < nw ? s_red[32 + lane] : 0.0, a2 = lane < nw ? s_red[64 + lane] : 0.0;
          a0 = warp_sum_d(a0); a1 = warp_sum_d(a1); a2 = warp_sum_d(a2);
          if (lane == 0) {
            FastRec* r = &rec[((ep1 + 1) & 1) * G + b];
            r->m = mb; r->s = a0; r->q = a1; r->sx = a2; r->nan = nan_b;
            __threadfence();
            st_release_u32(&r->epoch, ep1 + 1);
          }
        }
      }
      ep1++;
      // ---- B1: poll the G records, build the scaled table ----
      for (int j = tid; j < G; j += blockDim.x) {
        const FastRec* r = &rec[(ep1 & 1) * G + j];
        while (ld_acquire_u32(&r->epoch) != ep1) {}
        s_tab[j] = __ldcg(&r->m);
        s_tab[G + j] = __ldcg(&r->s);
        s_tab[2 * G + j] = __ldcg(&r->q);
        s_tab[3 * G + j] = __ldcg(&r->sx);
        if (__ldcg(&r->nan)) s_flag[0] = 1;
      }
      __syncthreads();
      // ---- P2: global max / sums / offsets: warp 0, fixed order => identical in every CTA ----
      if (wid == 0) {
        double M = NINF;
        for (int j = lane; j < G; j += 32) M = s_tab[j] > M ? s_tab[j] : M;
#pragma unroll
        for (int o = 16; o; o >>= 1) { double u = __shfl_xor_sync(0xffffffffu, M, o); M = u > M ? u : M; }
        double S = 0.0, Q = 0.0, SX = 0.0, carry = 0.0, mine_lo = 0.0, mine_hi = 0.0;
        for (int j0 = 0; j0 < G; j0 += 32) {
          int j = j0 + lane;
          double sc = 0.0, sj = 0.0, qj = 0.0, xj = 0.0;
          if (j < G) {
            sc = (s_tab[j] == NINF || M == NINF) ? 0.0 : exp(s_tab[j] - M);
            sj = s_tab[G + j] * sc; qj = s_tab[2 * G + j] * sc * sc; xj = s_tab[3 * G + j] * sc;
          }
          // inclusive scan of sj over the warp (fixed order)
          double inc = sj;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
          double incl = carry + inc;
          double prev = __shfl_up_sync(0xffffffffu, incl, 1);   // neighbour's inclusive value, bit-identical
          if (lane == 0) prev = carry;
          if (j == b) { mine_lo = prev; mine_hi = incl; }
          carry = __shfl_sync(0xffffffffu, incl, 31);
          Q += qj; SX += xj;
        }
        S = carry;
        Q = warp_sum_d(Q); SX = warp_sum_d(SX);
        mine_lo = warp_sum_d(mine_lo); mine_hi = warp_sum_d(mine_hi);  // only one lane is non-zero
        if (lane == 0) { s_misc[2] = M; s_misc[3] = S; s_misc[4] = Q; s_misc[5] = SX; s_misc[6] = mine_lo; s_misc[7] = mine_hi; }
      }
      __syncthreads();
      const double M = s_misc[2], S = s_misc[3], Q = s_misc[4], SX = s_misc[5];
      if (s_flag[0]) {  // NaN weight somewhere: R's `if (NA)` error
        if (b == 0 && tid == 0) { f.status[c] = 3; }
        dead = true;
      } else if (M < -1e8) {  // all(lw < -1e8): R/particle_filter_core.R:189-202
        loglike = NINF;
        if (b == 0 && tid == 0) { if (f.loglike_history) f.loglike_history[(size_t)c * f.T + obs] = NINF; f.early_exit[c] = 1; }
        dead = true;
      }
      if (dead) { __syncthreads(); if (tid == 0) s_flag[0] = 0; __syncthreads(); break; }
      loglike += (M + log(S) - log((double)n));
      const double ess = (S * S) / Q;
      int ralg = f.ralg;
      double thr = f.threshold;
      if (thr < 0) thr = (ralg == 0) ? __longlong_as_double(0x7FF0000000000000LL) : (ralg == 1 ? (double)n : (double)n / 2.0);
      const bool resample = (ralg == 0) ? false : (ralg == 1 ? true : (ess < thr));
      if (b == 0 && tid == 0) {
        if (f.loglike_history) f.loglike_history[(size_t)c * f.T + obs] = loglike;
        f.ess[(size_t)c * T1 + obs + 1] = resample ? (double)n : ess;
        if (!resample) f.state_est[(size_t)c * T1 + obs + 1] = SX / S;
      }
      if (!resample) continue;
      n_resampled++;

      // ---- P3: local cdf (normalised weights, fp64) and x into shared memory ----
      const double lo_cdf = s_misc[6] / S, hi_cdf = (b == G - 1) ? 2.0 : s_misc[7] / S;   // this CTA's cdf interval (lo, hi]
      const double wscale = ((mb == NINF) ? 0.0 : exp(mb - M)) / S;
      {
        double loc[FAST_PPT];
        double run = 0.0;
#pragma unroll
        for (int k = 0; k < FAST_PPT; k++) { run += (double)e[k] * wscale; loc[k] = run; }
        double inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_red[wid] = inc;
        __syncthreads();
        double woff = 0.0;
        for (int w = 0; w < wid; w++) woff += s_red[w];
        const double ex = lo_cdf + (woff + (inc - run));
#pragma unroll
        for (int k = 0; k < FAST_PPT; k++) {
          int li = tid * FAST_PPT + k;
          if (li < nb) { s_cdf[li] = ex + loc[k]; s_x[li] = x[k]; }
        }
      }
      // output range served by this CTA: positions in (lo_cdf, hi_cdf]
      PosGen pg;
      pg.key = key; pg.obs = (unsigned int)obs; pg.fn = P.resample_fn; pg.n = n; pg.pow2 = pow2; pg.inv_n = 1.0 / (double)n;
      pg.u_sys = 0.0;
      if (P.resample_fn == 1) { uint4x q0 = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, 0u); pg.u_sys = word_to_unit_f64(q0.w[0]); }
      if (tid == 0) s_flag[2] = pg.count_le(lo_cdf);
      if (tid == 32 % blockDim.x) s_flag[3] = (b == G - 1) ? n : pg.count_le(hi_cdf);
      __syncthreads();
      const int o_lo = s_flag[2], o_hi = s_flag[3];
      // ---- P4: search + scatter (outputs are handled in aligned quads: one Philox call per 4) ----
      double sumx = 0.0;
      if (n_loc > 0 && o_hi > o_lo) {
        Real fsum = 0;
        for (int qd = (o_lo >> 2) + tid; qd <= ((o_hi - 1) >> 2); qd += blockDim.x) {
          uint4x uq;
          if (P.resample_fn == 0) uq = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, (unsigned int)qd);
          int j = 0;
          Real v[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            int i = 4 * qd + k;
            v[k] = 0;
            if (i >= o_lo && i < o_hi) {
              double u = (P.resample_fn == 0) ? word_to_unit_f64(uq.w[k]) : pg.u_sys;
              double p = pg.pos(i, u);
              j = smem_lower_bound(s_cdf, j, n_loc, p);
              v[k] = s_x[j];
              fsum += v[k];
            }
          }
          int i0 = 4 * qd;
          if (i0 >= o_lo && i0 + 3 < o_hi) {
            if (sizeof(Real) == 4) __stcg((float4*)((float*)xnew + i0), make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]));
            else { __stcg((double2*)((double*)xnew + i0), make_double2((double)v[0], (double)v[1])); __stcg((double2*)((double*)xnew + i0 + 2), make_double2((double)v[2], (double)v[3])); }
          } else {
#pragma unroll
            for (int k = 0; k < 4; k++) if (i0 + k >= o_lo && i0 + k < o_hi) __stcg(&xnew[i0 + k], v[k]);
          }
        }
        sumx = (double)fsum;
      }
      // block sum of sumx, publish "done"
      {
        double v = warp_sum_d(sumx);
        __syncthreads();   // all scatter stores issued before the release below (and s_red reuse)
        if (lane == 0) s_red[wid] = v;
        __syncthreads();
        if (wid == 0) {
          double t = lane < nw ? s_red[lane] : 0.0;
          t = warp_sum_d(t);
          if (lane == 0) {
            FastRec2* r = &rec2[b];
            r->sumx = t;
            __threadfence();
            st_release_u32(&r->epoch, ep2 + 1);
          }
        }
      }
      ep2++;
      // ---- B2: wait for every CTA's scatter, reload this CTA's slice ----
      double tot = 0.0;
      for (int j = tid; j < G; j += blockDim.x) {
        const FastRec2* r = &rec2[j];
        while (ld_acquire_u32(&r->epoch) != ep2) {}
        tot += __ldcg(&r->sumx);
      }
      __syncthreads();
      if (b == 0) {
        double v = warp_sum_d(tot);
        if (lane == 0) s_red[wid] = v;
        __syncthreads();
        if (tid == 0) {
          double t = 0.0;
          for (int w = 0; w < nw; w++) t += s_red[w];
          f.state_est[(size_t)c * T1 + obs + 1] = t / (double)n;
        }
      }
      if (tid * FAST_PPT < nb) {
        const Real* src = xnew + ibase;
        if (sizeof(Real) == 4) {
          float4 a = __ldcg((const float4*)src), bb = __ldcg((const float4*)src + 1);
          x[0] = (Real)a.x; x[1] = (Real)a.y; x[2] = (Real)a.z; x[3] = (Real)a.w;
          x[4] = (Real)bb.x; x[5] = (Real)bb.y; x[6] = (Real)bb.z; x[7] = (Real)bb.w;
        } else {
#pragma unroll
          for (int k = 0; k < FAST_PPT; k += 2) { double2 a = __ldcg((const double2*)((const double*)src + k)); x[k] = (Real)a.x; x[k + 1] = (Real)a.y; }
        }
      }
    }  // obs
    if (b == 0 && tid == 0) { f.loglike[c] = loglike; f.n_resampled[c] = n_resampled; }
    __syncthreads();
  }    // filters
}

}  // namespace bssm
