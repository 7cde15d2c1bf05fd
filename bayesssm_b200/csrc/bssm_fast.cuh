// bssm_fast.cuh -- persistent bootstrap-filter kernel: the whole T loop of
// .particle_filter_core (R/particle_filter_core.R:123-246) for algorithm "BPF" in ONE launch.
//
// A filter is run by a GROUP of G co-resident CTAs (cooperative launch); CTA b owns the
// contiguous slice [b*nb, (b+1)*nb) of the particles and keeps it in REGISTERS for all T steps
// (PPT particles per thread).  Per observation:
//   P1  propagate (normals pre-generated while waiting at the previous sync point; one Philox call
//       per 4 particles) + log-weight; block max; e = exp(lw - max_b); block sums of e, e^2, e*x
//       and the block-local inclusive scan of e                              [registers + shuffles]
//   B1  every CTA publishes (max_b, sums) as an epoch-stamped record in L2, generates the NEXT
//       step's normals while the record travels, then polls the G records (release/acquire, no
//       atomics); all CTAs derive, redundantly but bit-identically, the global max / sum / ESS /
//       resampling decision and the cdf interval of every CTA                  [1 L2 round trip]
//   P3  INPUT-centric resampling: source j knows its cdf value c_j, hence -- in closed form -- the
//       number F(c_j) of output slots whose position (i + U_i)/N is <= c_j; it owns the output
//       slots [F(c_{j-1}), F(c_j)).  No search.  The stratified uniforms of the CTA's output range
//       are staged once in shared memory (one Philox call per 4 slots).
//   P4  the chosen x are scattered into a shared-memory staging buffer and copied out to x_new
//       with coalesced 16-byte stores
//   B2  every CTA reloads its slice of x_new; each element carries its epoch tag (LL protocol:
//       value and tag travel in one 8-byte word), so there is no fence, no flag and no second
//       round trip -- the reader simply polls its own elements               [< 1 L2 round trip]
// Records use the same self-validating words.  Only x_new (one write + one read per particle, L2
// resident) and the tiny records leave the SM.
// Same Philox keying and tie rule (first j with cdf[j] >= pos, clamp) as the general engine, so
// results do not depend on G or the launch geometry beyond floating-point summation order.
// In the throughput precision (Real = float) the within-thread part of the cdf and the slot
// arithmetic run in fp32 relative to an fp64 per-thread origin.
#pragma once
#include "bssm_common.cuh"
#include "bssm_filter.cuh"
#include "bssm_models.cuh"
#include "bssm_slots.cuh"

#include <type_traits>

namespace bssm {

constexpr int FAST_MAX_NB = 7168;    // particles per CTA
constexpr int FAST_MAX_G = 256;      // CTAs per group
constexpr int FAST_SLACK = 1024;     // staging capacity beyond the slice size
constexpr int FAST_HEAVY = 64;       // offspring count above which a source is expanded cooperatively
constexpr int FAST_HEAVY_CAP = 64;
// output slots per thread of one expansion pass: 25 % beyond the slice, rounded up to whole 16-byte accesses
__host__ __device__ constexpr int fast_spt(int ppt) { return (ppt * 5 / 4 + 3) & ~3; }

// LL ("low latency") words: 32 data bits + 32-bit epoch tag in one 8-byte unit, two units per
// 16-byte access.  A reader that sees the expected tag also sees the data: no fence, no separate
// flag, no dependent second load.  The accesses are relaxed at GPU scope (all that an exchange between
// CTAs of one GPU needs): +4 % over `volatile`, which is system scope (profiles/r1_ab_experiments.md).
struct __align__(128) FastRec {   // published once per observation by each CTA: 5 doubles as (lo, tag, hi, tag)
  uint4 w[5];                     // m, s, q, sx, pending sum of the previous step's resampled x
  uint4 pad[3];
};

struct FastParams {
  FilterDev f;
  int G, ngroups;
  int resample_fn;
  FastRec* rec;     // [ngroups][2][G]
  void* xnew;       // [ngroups][G * nb_max] LL elements: uint2 (f32) / uint4 (f64)
  int nb_max;       // slice stride (multiple of PPT)
  int cap;          // staging capacity (outputs) = nb_max + FAST_SLACK
  long long* timing;  // optional [gridDim][16] phase cycle counters (BSSM_FAST_TIMING=1, diagnostics)
};

#ifndef BSSM_EMU
__device__ __forceinline__ void ll_store_v4(void* p, unsigned int a, unsigned int b, unsigned int c, unsigned int d) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ll_store_v2(void* p, unsigned int a, unsigned int b) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint4 ll_load_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
#else   // CPU logic test (tests/simt_emu.h): each 8-byte (value, tag) unit is one atomic access; a poll lets the others run
inline void ll_store_v4(void* p, unsigned int a, unsigned int b, unsigned int c, unsigned int d) {
  __atomic_store_n((unsigned long long*)p, ((unsigned long long)b << 32) | a, __ATOMIC_RELEASE);
  __atomic_store_n((unsigned long long*)p + 1, ((unsigned long long)d << 32) | c, __ATOMIC_RELEASE);
}
inline void ll_store_v2(void* p, unsigned int a, unsigned int b) { __atomic_store_n((unsigned long long*)p, ((unsigned long long)b << 32) | a, __ATOMIC_RELEASE); }
inline uint4 ll_load_v4(const void* p) {
  emu_poll_yield();
  const unsigned long long lo = __atomic_load_n((const unsigned long long*)p, __ATOMIC_ACQUIRE), hi = __atomic_load_n((const unsigned long long*)p + 1, __ATOMIC_ACQUIRE);
  return make_uint4((unsigned int)lo, (unsigned int)(lo >> 32), (unsigned int)hi, (unsigned int)(hi >> 32));
}
#endif
__device__ __forceinline__ void ll_put_double(uint4* p, double d, unsigned int tag) {
  unsigned long long b = (unsigned long long)__double_as_longlong(d);
  ll_store_v4(p, (unsigned int)b, tag, (unsigned int)(b >> 32), tag);
}
__device__ __forceinline__ double ll_get_double(const uint4& v) {
  return __longlong_as_double((long long)(((unsigned long long)v.z << 32) | v.x));
}
// publish a 5-double record / poll one until every word carries `tag`
__device__ __forceinline__ void rec_publish(FastRec* r, double m, double s, double q, double sx, double sp, unsigned int tag) {
  ll_put_double(&r->w[0], m, tag); ll_put_double(&r->w[1], s, tag); ll_put_double(&r->w[2], q, tag);
  ll_put_double(&r->w[3], sx, tag); ll_put_double(&r->w[4], sp, tag);
}
__device__ __forceinline__ void rec_poll(const FastRec* r, unsigned int tag, double* out /*5*/) {
  uint4 v[5];
  bool ok;
  do {
    ok = true;
#pragma unroll
    for (int i = 0; i < 5; i++) { v[i] = ll_load_v4(&r->w[i]); ok = ok && v[i].y == tag && v[i].w == tag; }
  } while (!ok);
#pragma unroll
  for (int i = 0; i < 5; i++) out[i] = ll_get_double(v[i]);
}
#ifndef BSSM_EMU
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

template <typename Model, typename Real, int PPT, bool HEADS>
__global__ void __launch_bounds__(FAST_MAX_NB / PPT, 1) k_fast_bpf(FastParams P) {
  static_assert(Model::D == 1 && Model::NZ_TRANS == 1 && Model::NU_TRANS == 0 && Model::NZ_INIT == 1 && Model::NU_INIT == 0,
                "persistent kernel: 1-D models with one normal per transition");
  static_assert(PPT % 4 == 0, "one Philox call serves 4 particles");
  constexpr bool F32 = sizeof(Real) == 4;
#ifndef BSSM_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#else
  unsigned char* const smem_raw = emu_dynamic_smem();
#endif
  const FilterDev& f = P.f;
  const int G = P.G;
  const int group = blockIdx.x / G, b = blockIdx.x % G;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = (blockDim.x + 31) >> 5;
  // shared memory carve-up
  double* s_tab = (double*)smem_raw;                 // [5][G]: m, s -> inclusive cdf numerator A, q, sx, pending of every CTA
  double* s_red = s_tab + ((5 * G + 1) & ~1);        // [5][32] per-warp partials (16-byte aligned)
  Real* s_out = (Real*)(s_red + 5 * 32);             // [cap] staging of the chosen x
  unsigned int* s_u = (unsigned int*)(s_out + P.cap);  // [cap] staged stratified uniforms (raw words)
  unsigned int* s_head = s_u + P.cap;                // [cap] (HEADS only) expansion: (source index << 16 | index of its x in s_x) at the first slot of a source
  Real* s_x = (Real*)(s_head + P.cap);               // [blockDim.x * PPT] this CTA's particles, [PPT/4][threads] x 16 B (conflict-free)
  constexpr int SPT = fast_spt(PPT);                 // output slots per thread in the expansion: cap = blockDim.x * SPT
  static_assert(SPT % 4 == 0, "16-byte accesses to the head / staging arrays");
  __shared__ int s_wf[32];
  __shared__ unsigned int s_wh[32];
  __shared__ int s_heavy_n;
  __shared__ int s_heavy_lo[FAST_HEAVY_CAP], s_heavy_hi[FAST_HEAVY_CAP];
  __shared__ Real s_heavy_x[FAST_HEAVY_CAP];
  __shared__ double s_pending;                       // sum of the x chosen by this CTA in the last resampling

  FastRec* rec = P.rec + (size_t)group * 2 * G;
  typedef typename std::conditional<F32, uint2, uint4>::type XEl;   // LL element of x_new
  XEl* xnew = (XEl*)P.xnew + (size_t)group * G * P.nb_max;
  unsigned int ep1 = 0, ep2 = 0;   // record / x_new epochs: identical sequences in every CTA of the group
  const double NINF = -__longlong_as_double(0x7FF0000000000000LL);
  // optional phase timing (diagnostics): compiled in only with -DBSSM_FAST_TIMING_BUILD, because the counters
  // would otherwise hold ~26 registers for the whole kernel
#ifdef BSSM_FAST_TIMING_BUILD
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define FAST_TICK(ph) do { if (P.timing && tid == 0) { long long t_ = clock64(); tacc[ph] += t_ - tprev; tprev = t_; } } while (0)
#else
#define FAST_TICK(ph) do { } while (0)
#endif
  const int R = (G + 31) >> 5;   // records per lane in P2

  for (int c = group; c < f.C; c += P.ngroups) {
    if (!f.alive[c]) continue;
    const int n = filt_n(f, c);
    int nb = (n + G - 1) / G;
    nb = (nb + PPT - 1) / PPT * PPT;
    const int base = b * nb;                                  // first global particle of this CTA
    const int n_loc = max(0, min(n - base, nb));              // particles owned by this CTA
    const int ibase = base + tid * PPT;                       // first global particle of this thread
    const int n_own = max(0, min(PPT, min(n - ibase, nb - tid * PPT)));  // owned particles of this thread
    Real par[Model::NPAR];
    Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
    const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
    const int T1 = f.T + 1;
    const int ralg = f.ralg;
    double thr = f.threshold;
    if (thr < 0) thr = (ralg == 0) ? __longlong_as_double(0x7FF0000000000000LL) : (ralg == 1 ? (double)n : (double)n / 2.0);
    const double log_n = log((double)n);

    // ---- init (R/particle_filter_core.R:76-116) ----
    Real x[PPT];
    double sum0 = 0.0;
#pragma unroll
    for (int h = 0; h < PPT / 4; h++) {
      uint4x qd = noise_quad(key, T_INIT, TAG_INIT_Z, 0u, (unsigned int)(ibase + 4 * h) >> 2);
      Real zz[4];
      Math<Real>::box_muller(qd.w[0], qd.w[1], zz[0], zz[1]);
      Math<Real>::box_muller(qd.w[2], qd.w[3], zz[2], zz[3]);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        Real xi[1]; Real zi[1] = {zz[k]};
        Model::template init<Real>(xi, par, zi, nullptr);
        x[4 * h + k] = (4 * h + k < n_own) ? xi[0] : (Real)0;   // padding lanes stay finite
        sum0 += (double)x[4 * h + k];
      }
    }
    int n_resampled = 0;
    double loglike = 0.0;   // meaningful in thread 0 of CTA 0
    if (tid == 0) { s_heavy_n = 0; s_pending = 0.0; }
    if constexpr (HEADS) {
#pragma unroll
      for (int i = 0; i < SPT; i++) s_head[tid * SPT + i] = 0u;   // every thread keeps its own slots of the head array clear
    }
    int pending_obs = -1;   // observation whose resampled state estimate is still to be written
    // t = 0 state estimate: block sum -> record; CTA 0 gathers
    {
      double v = warp_sum_d(sum0);
      if (lane == 0) s_red[wid] = v;
      __syncthreads();
      if (wid == 0) {
        double t = lane < nw ? s_red[lane] : 0.0;
        t = warp_sum_d(t);
        if (lane == 0) rec_publish(&rec[((ep1 + 1) & 1) * G + b], 0.0, 0.0, 0.0, t, 0.0, ep1 + 1);
      }
      ep1++;
      // every CTA polls every record (not only CTA 0): a record buffer may only be reused once all CTAs
      // have passed the poll of the epoch before, which is what orders the two-deep record buffers
      {
        double v0 = 0.0;
        for (int j = tid; j < G; j += blockDim.x) {
          double rv[5];
          rec_poll(&rec[(ep1 & 1) * G + j], ep1, rv);
          v0 += rv[3];
        }
        v0 = warp_sum_d(v0);
        __syncthreads();
        if (lane == 0) s_red[wid] = v0;
        __syncthreads();
        if (b == 0 && tid == 0) {
          double t = 0.0;
          for (int w = 0; w < nw; w++) t += s_red[w];
          f.ess[(size_t)c * T1] = (double)n;
          f.state_est[(size_t)c * T1] = t / (double)n;
        }
      }
      __syncthreads();
    }

    // normals of the next transition, generated ahead of time (they do not depend on x)
    Real zpre[PPT];
    int zpre_t = -1;   // absolute time index (tnow - 1) the pre-generated normals belong to
    auto gen_normals = [&](int tz) {
#pragma unroll
      for (int h = 0; h < PPT / 4; h++) {
        uint4x qd = noise_quad(key, (unsigned int)tz, TAG_TRANS_Z, 0u, (unsigned int)(ibase + 4 * h) >> 2);
        Math<Real>::box_muller(qd.w[0], qd.w[1], zpre[4 * h + 0], zpre[4 * h + 1]);
        Math<Real>::box_muller(qd.w[2], qd.w[3], zpre[4 * h + 2], zpre[4 * h + 3]);
      }
      zpre_t = tz;
    };

    double ynext[4] = {0, 0, 0, 0};
    if (f.T > 0) for (int k = 0; k < f.dy && k < 4; k++) ynext[k] = f.y[k];
    for (int obs = 0; obs < f.T; obs++) {
      const int ot = f.obs_times ? f.obs_times[obs] : obs + 1;
      const int prev_t = obs == 0 ? 0 : (f.obs_times ? f.obs_times[obs - 1] : obs);
      double yv[4] = {ynext[0], ynext[1], ynext[2], ynext[3]};
      if (obs + 1 < f.T) for (int k = 0; k < f.dy && k < 4; k++) ynext[k] = f.y[(size_t)(obs + 1) * f.dy + k];   // prefetch

      FAST_TICK(0);
      // ---- P1: propagate + log-weight ----
      for (int tnow = prev_t + 1; tnow <= ot; tnow++) {
        if (zpre_t != tnow - 1) gen_normals(tnow - 1);
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          Real zi[1] = {zpre[k]};
          Model::template transition<Real>(&x[k], par, tnow, zi, nullptr);
        }
      }
      Real e[PPT];   // first the log-weights, then exp(lw - block max)
      Real mloc = Math<Real>::ninf();
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        e[k] = Model::template loglik<Real>(yv, &x[k], par, ot);
        if (k >= n_own) e[k] = Math<Real>::ninf();
        mloc = e[k] > mloc ? e[k] : mloc;
      }
      // block max: per-warp partials, then every warp reduces the partials redundantly (one barrier)
      {
        Real v = mloc;
#pragma unroll
        for (int o = 16; o; o >>= 1) { Real t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
        if (lane == 0) s_red[wid] = (double)v;
      }
      __syncthreads();
      const double mb = warp_max_d(lane < nw ? s_red[lane] : NINF);
      // e = exp(lw - mb); thread sums; block-local inclusive scan of e
      double exu;            // block-local exclusive prefix of this thread (unnormalised)
      {
        Real fs = 0, fq = 0, fx = 0;
        const Real mbr = (mb == NINF) ? (Real)0 : (Real)mb;   // exp(-inf - 0) = 0: no per-particle guard
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          Real ek = Math<Real>::exp_(e[k] - mbr);
          e[k] = ek;
          fs += ek; fq += ek * ek; fx += ek * x[k];
        }
        const double run = (double)fs;
        double inc = warp_incl_scan_d(run, lane);
        double tq = warp_sum_d((double)fq), tx = warp_sum_d((double)fx);
        __syncthreads();   // s_red partials of the max have been consumed by every warp
        if (lane == 31) s_red[wid] = inc;            // warp totals of e
        if (lane == 0) { s_red[32 + wid] = tq; s_red[64 + wid] = tx; }
        double prev = __shfl_up_sync(0xffffffffu, inc, 1);
        exu = lane == 0 ? 0.0 : prev;                // exclusive within the warp
        __syncthreads();
        // every warp scans the warp totals redundantly
        double wt = lane < nw ? s_red[lane] : 0.0;
        double winc = warp_incl_scan_d(wt, lane);
        exu += __shfl_sync(0xffffffffu, winc - wt, wid);   // + exclusive prefix of this warp
        if (wid == 0) {
          double a0 = __shfl_sync(0xffffffffu, winc, 31);
          double a1 = warp_sum_d(lane < nw ? s_red[32 + lane] : 0.0), a2 = warp_sum_d(lane < nw ? s_red[64 + lane] : 0.0);
          if (lane == 0) rec_publish(&rec[((ep1 + 1) & 1) * G + b], mb, a0, a1, a2, s_pending, ep1 + 1);
        }
      }
      ep1++;
      FAST_TICK(1);   // P1 (propagate, weights, block reductions, publish)
      // overlap the L2 round trip with the next observation's normals
      if (obs + 1 < f.T) gen_normals(ot);
      FAST_TICK(2);   // next-step normals
      // ---- B1: poll the G records ----
      for (int j = tid; j < G; j += blockDim.x) {
        double rv[5];
        rec_poll(&rec[(ep1 & 1) * G + j], ep1, rv);
        s_tab[j] = rv[0]; s_tab[G + j] = rv[1]; s_tab[2 * G + j] = rv[2]; s_tab[3 * G + j] = rv[3]; s_tab[4 * G + j] = rv[4];
      }
      __syncthreads();
      FAST_TICK(3);   // B1 poll + barrier
      // ---- P2: global max / sums / this CTA's cdf interval.  EVERY warp evaluates the same expressions on the
      //      same table (R consecutive records per lane, one warp scan), so the results are warp-uniform
      //      registers, bit-identical in every warp of every CTA: no roles, no broadcast, no barrier ----
      int dead = 0, resample = 0;
      double lo_cdf = 0.0, hi_cdf = 0.0, wscale = 0.0;
      {
        double M = NINF;
        for (int j = lane; j < G; j += 32) M = s_tab[j] > M ? s_tab[j] : M;
        M = warp_max_d(M);
        const int j0 = lane * R;
        double loc_s = 0.0, loc_q = 0.0, loc_x = 0.0, loc_p = 0.0, my_lo = 0.0, my_hi = 0.0;
        for (int r = 0; r < R; r++) {
          const int j = j0 + r;
          if (j < G) {
            const double mj = s_tab[j];
            double sc = 0.0;
            if (!(mj == NINF || M == NINF)) sc = F32 ? (double)__expf((float)(mj - M)) : exp(mj - M);
            loc_s += s_tab[G + j] * sc;
            if (j == b - 1) my_lo = loc_s;            // thread-local inclusive values of records b-1 and b
            if (j == b) my_hi = loc_s;
            loc_q += s_tab[2 * G + j] * sc * sc; loc_x += s_tab[3 * G + j] * sc; loc_p += s_tab[4 * G + j];
          }
        }
        const double inc = warp_incl_scan_d(loc_s, lane);
        const double off = inc - loc_s;               // everything before this lane's first record
        const double S = __shfl_sync(0xffffffffu, inc, 31);
        const double Q = warp_sum_d(loc_q), SX = warp_sum_d(loc_x), PEND = warp_sum_d(loc_p);
        const double A_hi = __shfl_sync(0xffffffffu, off + my_hi, b / R);
        const double A_lo = b == 0 ? 0.0 : __shfl_sync(0xffffffffu, off + my_lo, (b - 1) / R);
        const bool bad = (S != S) || (SX != SX) || (M != M);
        const bool empty = M < -1e8;
        dead = (bad || empty) ? 1 : 0;
        // ess < thr  <=>  S^2 < thr * Q  (no division on the critical path)
        resample = dead ? 0 : ((ralg == 0) ? 0 : (ralg == 1 ? 1 : (S * S < thr * Q)));
        if (resample) {
          const double invS = 1.0 / S;
          lo_cdf = A_lo * invS;
          hi_cdf = (b == G - 1) ? 2.0 : A_hi * invS;
          double w = 0.0;
          if (mb != NINF) w = F32 ? (double)__expf((float)(mb - M)) : exp(mb - M);
          wscale = w * invS;
        }
        if (b == 0 && tid == 0) {
          // running log-likelihood and the outputs of this observation: one thread, off everybody's critical path
          if (pending_obs >= 0) f.state_est[(size_t)c * T1 + pending_obs + 1] = PEND / (double)n;
          if (bad) {                       // NaN weight somewhere: R's `if (NA)` error
            f.status[c] = 3;
          } else if (empty) {              // all(lw < -1e8): R/particle_filter_core.R:189-202
            loglike = NINF;
            if (f.loglike_history) f.loglike_history[(size_t)c * f.T + obs] = NINF;
            f.early_exit[c] = 1;
          } else {
            loglike += (M + log(S) - log_n);
            if (f.loglike_history) f.loglike_history[(size_t)c * f.T + obs] = loglike;
            f.ess[(size_t)c * T1 + obs + 1] = resample ? (double)n : (S * S) / Q;
            if (!resample) f.state_est[(size_t)c * T1 + obs + 1] = SX / S;
          }
        }
      }
      FAST_TICK(4);   // P2
      pending_obs = -1;
      if (dead) break;
      if (!resample) continue;
      n_resampled++;
      pending_obs = obs;

      // ---- P3: closed-form offspring ranges ----
      SlotCounter sc;
      sc.key = key; sc.obs = (unsigned int)obs; sc.fn = P.resample_fn; sc.n = n; sc.w_sys = 0u;
      sc.s_u = s_u; sc.u_cap = P.cap;
      {
        double t0 = lo_cdf * (double)n;
        int i0 = t0 >= (double)n ? n : (int)t0;
        sc.u_base = max(0, (i0 & ~3) - 4);
      }
      if (P.resample_fn == 1) {
        uint4x q0 = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, 0u);
        sc.w_sys = q0.w[0];
      } else {
        // stage the Philox words of the slots this CTA is expected to serve: one call per 4 slots
        const int q_end = min((n + 3) >> 2, (sc.u_base + P.cap) >> 2);
        for (int qd = (sc.u_base >> 2) + tid; qd < q_end; qd += blockDim.x) {
          uint4x uq = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, (unsigned int)qd);
          *(uint4*)&s_u[4 * qd - sc.u_base] = make_uint4(uq.w[0], uq.w[1], uq.w[2], uq.w[3]);
        }
        __syncthreads();
      }
      FAST_TICK(5);   // stage uniforms
      const int o_lo = sc.count_le(lo_cdf);
      const int o_hi = (b == G - 1) ? n : sc.count_le(hi_cdf);
      // F of this thread's sources (monotone by a running max; clamped into [o_lo, o_hi])
      int F[PPT];
      {
        int fmax = o_lo;
        if (F32) {
          // fp64 origin per thread, fp32 increments: t_k = T0 + (sum of e up to k) * wscale * n
          const double T0 = (lo_cdf + exu * wscale) * (double)n;
          const double T0c = T0 < (double)n ? T0 : (double)n;
          const int I0 = (int)T0c;
          const float f0 = (float)(T0c - (double)I0);
          const float wsn = (float)(wscale * (double)n);
          float accf = 0.f;
#pragma unroll
          for (int k = 0; k < PPT; k++) {
            accf += (float)e[k];
            const float tf = fmaf(accf, wsn, f0);
            // floor and fraction without conversion instructions (1.5 * 2^23 trick)
            const float r = (tf - 0.5f) + 12582912.0f;
            const int ii = __float_as_int(r) - 0x4B400000;
            const float frac = tf - (r - 12582912.0f);                    // in [0, 1]
            const float g = frac + 1.0f;                                   // [1, 2]
            const unsigned int fbits = ((unsigned int)__float_as_int(g) & 0x7FFFFFu) << 9;   // frac * 2^32, 23 bits
            const int i = I0 + ii;
            int v;
            if (i >= n) v = n;
            else v = i + ((sc.word_of(i) < fbits || g >= 2.0f) ? 1 : 0);   // (i + U_i) <= t  <=>  U_i <= frac
            if (tid * PPT + k == n_loc - 1) v = o_hi;     // clamp: the last particle takes what is left
            v = min(max(v, o_lo), o_hi);
            if (k >= n_own) v = o_lo;
            fmax = max(fmax, v);
            F[k] = fmax;
          }
        } else {
          double acc = exu;
#pragma unroll
          for (int k = 0; k < PPT; k++) {
            acc += (double)e[k];
            int v = o_lo;
            if (k < n_own) {
              v = sc.count_le(lo_cdf + acc * wscale);
              if (tid * PPT + k == n_loc - 1) v = o_hi;
              v = min(max(v, o_lo), o_hi);
            }
            fmax = max(fmax, v);
            F[k] = fmax;
          }
        }
      }
      // exclusive prefix-max of the per-thread last F over the block
      int prevF;
      {
        int inc = F[PPT - 1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, t); }
        if (lane == 31) s_wf[wid] = inc;
        prevF = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) prevF = o_lo;
        __syncthreads();
        int wv = lane < nw ? s_wf[lane] : o_lo;
        int winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc = max(winc, t); }
        int wprev = __shfl_sync(0xffffffffu, winc, (wid + 31) & 31);
        if (wid == 0) wprev = o_lo;
        prevF = max(prevF, wprev);
#pragma unroll
        for (int k = 0; k < PPT; k++) F[k] = max(F[k], prevF);
      }
      FAST_TICK(6);   // offspring ranges + prefix max
      // ---- P4: expansion, output-centric (chunks of `cap` slots).  Every source with offspring in the chunk marks the
      //      first of its slots with (source index, index of its x in s_x); a running maximum over the slots -- source
      //      indices grow with the slot -- tells every slot its source: O(1) per source and per slot, no loop over
      //      the offspring of a source, no special case for heavy sources.  The chosen x are staged and leave the
      //      SM as coalesced 16-byte stores ----
      //      HEADS = false (big slices, 16 particles per thread): per-source scatter loops into the staging buffer --
      //      measured 2 % faster there (the expansion pays two more barriers), 20 % slower on small slices ----
      {
        Real sumx = 0;
        if constexpr (HEADS) {
        const int o_base = o_lo & ~3;
        const int nthr = blockDim.x;
        // this CTA's particles into shared memory (the registers are reloaded from x_new below)
#pragma unroll
        for (int h4 = 0; h4 < PPT / 4; h4++) {
          if (F32) *(float4*)&s_x[(h4 * nthr + tid) * 4] = make_float4((float)x[4 * h4], (float)x[4 * h4 + 1], (float)x[4 * h4 + 2], (float)x[4 * h4 + 3]);
          else { s_x[(h4 * nthr + tid) * 4] = x[4 * h4]; s_x[(h4 * nthr + tid) * 4 + 1] = x[4 * h4 + 1]; s_x[(h4 * nthr + tid) * 4 + 2] = x[4 * h4 + 2]; s_x[(h4 * nthr + tid) * 4 + 3] = x[4 * h4 + 3]; }
        }
        for (int c0 = o_base; c0 < o_hi; c0 += P.cap) {
          const int c1 = min(o_hi, c0 + P.cap);
          int lo_k = prevF;
#pragma unroll
          for (int k = 0; k < PPT; k++) {
            const int hi_k = F[k];
            const int a = max(lo_k, c0);
            if (min(hi_k, c1) > a) s_head[a - c0] = ((unsigned int)(tid * PPT + k) << 16) | (unsigned int)(((k >> 2) * nthr + tid) * 4 + (k & 3));
            if (c0 == o_base && hi_k > lo_k) {
              // count as a float without a conversion instruction (exact below 2^23)
              const float cf = __int_as_float(0x4B000000 | (hi_k - lo_k)) - 8388608.0f;
              sumx += (Real)cf * x[k];
            }
            lo_k = max(lo_k, hi_k);
          }
          __syncthreads();
          unsigned int hd[SPT];
#pragma unroll
          for (int i = 0; i < SPT; i += 4) { const uint4 v4 = *(const uint4*)&s_head[tid * SPT + i]; hd[i] = v4.x; hd[i + 1] = v4.y; hd[i + 2] = v4.z; hd[i + 3] = v4.w; }
#pragma unroll
          for (int i = 0; i < SPT; i += 4) *(uint4*)&s_head[tid * SPT + i] = make_uint4(0u, 0u, 0u, 0u);   // clear for the next chunk / step
#pragma unroll
          for (int i = 1; i < SPT; i++) hd[i] = max(hd[i], hd[i - 1]);
          unsigned int inc = hd[SPT - 1];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, t); }
          if (lane == 31) s_wh[wid] = inc;
          unsigned int carry = __shfl_up_sync(0xffffffffu, inc, 1);
          if (lane == 0) carry = 0u;
          __syncthreads();
          {
            unsigned int wv = lane < nw ? s_wh[lane] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, wv, o); if (lane >= o) wv = max(wv, t); }
            const unsigned int wprev = __shfl_sync(0xffffffffu, wv, (wid + 31) & 31);
            if (wid > 0) carry = max(carry, wprev);
          }
          Real val[SPT];
#pragma unroll
          for (int i = 0; i < SPT; i++) val[i] = s_x[max(hd[i], carry) & 0xFFFFu];
#pragma unroll
          for (int i = 0; i < SPT; i += 4) {
            if (F32) *(float4*)&s_out[tid * SPT + i] = make_float4((float)val[i], (float)val[i + 1], (float)val[i + 2], (float)val[i + 3]);
            else { s_out[tid * SPT + i] = val[i]; s_out[tid * SPT + i + 1] = val[i + 1]; s_out[tid * SPT + i + 2] = val[i + 2]; s_out[tid * SPT + i + 3] = val[i + 3]; }
          }
          __syncthreads();
          // copy out as LL elements (value + epoch tag): 16-byte stores, 8-byte at the ragged ends
          const int first = max(c0, o_lo), last = c1;   // slots [first, last) are valid in this chunk
          const unsigned int tag = ep2 + 1;
          if (F32) {
            for (int o = c0 + 2 * tid; o < last; o += 2 * blockDim.x) {
              const float v0 = (float)s_out[o - c0], v1 = (float)s_out[o + 1 - c0];
              if (o >= first && o + 1 < last) ll_store_v4(&xnew[o], __float_as_uint(v0), tag, __float_as_uint(v1), tag);
              else {
                if (o >= first && o < last) ll_store_v2(&xnew[o], __float_as_uint(v0), tag);
                if (o + 1 >= first && o + 1 < last) ll_store_v2(&xnew[o + 1], __float_as_uint(v1), tag);
              }
            }
          } else {
            for (int o = max(first, c0) + tid; o < last; o += blockDim.x) ll_put_double((uint4*)&xnew[o], (double)s_out[o - c0], tag);
          }
          __syncthreads();
        }
        } else {
        const int o_base = o_lo & ~3;
        for (int c0 = o_base; c0 < o_hi; c0 += P.cap) {
          const int c1 = min(o_hi, c0 + P.cap);
          int lo_k = prevF;
#pragma unroll
          for (int k = 0; k < PPT; k++) {
            const int hi_k = F[k];
            const int a = max(lo_k, c0);
            int cnt = min(hi_k, c1) - a;
            if (c0 == o_base && hi_k > lo_k) {
              // count as a float without a conversion instruction (exact below 2^23)
              const float cf = __int_as_float(0x4B000000 | (hi_k - lo_k)) - 8388608.0f;
              sumx += (Real)cf * x[k];
            }
            if (cnt > FAST_HEAVY) {
              int slot = atomicAdd(&s_heavy_n, 1);
              if (slot < FAST_HEAVY_CAP) { s_heavy_lo[slot] = a; s_heavy_hi[slot] = a + cnt; s_heavy_x[slot] = x[k]; cnt = 0; }
            }
            // warp-uniform trip count: no divergent loop bookkeeping
            const int mx = __reduce_max_sync(0xffffffffu, cnt);
            Real* dst = s_out + (a - c0);
            for (int r = 0; r < mx; r++) if (r < cnt) dst[r] = x[k];
            lo_k = max(lo_k, hi_k);
          }
          __syncthreads();
          const int nh = min(s_heavy_n, FAST_HEAVY_CAP);
          for (int h = 0; h < nh; h++) {
            const int a = s_heavy_lo[h], z = s_heavy_hi[h];
            const Real xv = s_heavy_x[h];
            for (int o = a + tid; o < z; o += blockDim.x) s_out[o - c0] = xv;
          }
          if (nh) __syncthreads();
          // copy out as LL elements (value + epoch tag): 16-byte stores, 8-byte at the ragged ends
          const int first = max(c0, o_lo), last = c1;   // slots [first, last) are valid in this chunk
          const unsigned int tag = ep2 + 1;
          if (F32) {
            for (int o = c0 + 2 * tid; o < last; o += 2 * blockDim.x) {
              const float v0 = (float)s_out[o - c0], v1 = (float)s_out[o + 1 - c0];
              if (o >= first && o + 1 < last) ll_store_v4(&xnew[o], __float_as_uint(v0), tag, __float_as_uint(v1), tag);
              else {
                if (o >= first && o < last) ll_store_v2(&xnew[o], __float_as_uint(v0), tag);
                if (o + 1 >= first && o + 1 < last) ll_store_v2(&xnew[o + 1], __float_as_uint(v1), tag);
              }
            }
          } else {
            for (int o = max(first, c0) + tid; o < last; o += blockDim.x) ll_put_double((uint4*)&xnew[o], (double)s_out[o - c0], tag);
          }
          if (tid == 0) s_heavy_n = 0;
          __syncthreads();
        }
        }
        // block sum of the chosen x: travels in the next record (state estimate after resampling)
        double v = warp_sum_d((double)sumx);
        if (lane == 0) s_red[wid] = v;
        __syncthreads();
        if (wid == 0) {
          double t = lane < nw ? s_red[lane] : 0.0;
          t = warp_sum_d(t);
          if (lane == 0) s_pending = t;
        }
      }
      ep2++;
      FAST_TICK(7);   // scatter + copy-out + block sum
      // ---- B2: poll this thread's own elements of x_new until they carry this step's tag ----
      if (n_own > 0) {
        const XEl* src = xnew + ibase;
        bool ok;
        if (F32) {
          do {
            ok = true;
#pragma unroll
            for (int h = 0; h < PPT / 2; h++) {
              const uint4 v = ll_load_v4(src + 2 * h);
              ok = ok && (v.y == ep2 || 2 * h >= n_own) && (v.w == ep2 || 2 * h + 1 >= n_own);
              x[2 * h] = (Real)__uint_as_float(v.x); x[2 * h + 1] = (Real)__uint_as_float(v.z);
            }
          } while (!ok);
        } else {
          do {
            ok = true;
#pragma unroll
            for (int k = 0; k < PPT; k++) {
              const uint4 v = ll_load_v4(src + k);
              ok = ok && ((v.y == ep2 && v.w == ep2) || k >= n_own);
              x[k] = (Real)ll_get_double(v);
            }
          } while (!ok);
        }
        if (n_own < PPT) {
#pragma unroll
          for (int k = 0; k < PPT; k++) if (k >= n_own) x[k] = (Real)0;   // x_new beyond n is never written
        }
      }
      FAST_TICK(8);   // B2 poll (reload) -- no barrier here: warps whose elements arrived start the next step
    }  // obs
    // flush: the state estimate of a final resampling step still travels in the records
    __syncthreads();
    if (tid == 0) rec_publish(&rec[((ep1 + 1) & 1) * G + b], 0.0, 0.0, 0.0, 0.0, s_pending, ep1 + 1);
    ep1++;
    {
      double v0 = 0.0;
      for (int j = tid; j < G; j += blockDim.x) {   // all CTAs poll: see the note at the t = 0 exchange
        double rv[5];
        rec_poll(&rec[(ep1 & 1) * G + j], ep1, rv);
        v0 += rv[4];
      }
      v0 = warp_sum_d(v0);
      if (lane == 0) s_red[wid] = v0;
      __syncthreads();
      if (b == 0 && tid == 0) {
        double t = 0.0;
        for (int w = 0; w < nw; w++) t += s_red[w];
        if (pending_obs >= 0) f.state_est[(size_t)c * T1 + pending_obs + 1] = t / (double)n;
        f.loglike[c] = loglike; f.n_resampled[c] = n_resampled;
      }
    }
    __syncthreads();
  }    // filters
#ifdef BSSM_FAST_TIMING_BUILD
  if (P.timing && tid == 0) for (int i = 0; i < 12; i++) P.timing[(size_t)blockIdx.x * 16 + i] = tacc[i];
#endif
#undef FAST_TICK
}

}  // namespace bssm
