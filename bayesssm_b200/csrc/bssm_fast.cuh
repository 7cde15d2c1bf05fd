// bssm_fast.cuh -- persistent bootstrap-filter kernel: the whole T loop of
// .particle_filter_core (R/particle_filter_core.R:123-246) for algorithm "BPF" in ONE launch.
//
// A filter is run by a GROUP of G co-resident CTAs (cooperative launch); CTA b owns the contiguous slice
// [b*nb, (b+1)*nb) of the particles and keeps it in REGISTERS for all T steps (PPT particles per thread).
// A CTA is NW WORKER warps plus one SERVICE warp; the workers never meet at a block barrier of their own.
// Per observation:
//   P1  (workers) propagate (normals pre-generated during the previous wait; one Philox call per 4
//       particles) + log-weight; WARP max; e = exp(lw - max_w); warp sums of e, e^2, e*x and the
//       warp-local inclusive scan of e.  One 5-double record per warp goes to shared memory;
//       the warp ARRIVES at named barrier A and goes on with work that does not depend on the exchange:
//       the next observation's normals and the stratified uniforms of the slots it will probably
//       serve (a window around its own slice; misses are repaired per warp, never wrong)
//   B1  (service warp) waits on A, folds the warp records into the CTA record (max rescale, warp
//       prefix), publishes it as epoch-stamped LL words in L2, polls the G records of the group, and
//       derives -- bit-identically in every CTA -- the global max / sum / ESS / resampling decision, the
//       cdf interval of the CTA and, per worker warp, the slot position of its first particle and
//       its slots-per-unit-weight scale.  It arrives at named barrier B, where the workers wait
//                                                                              [1 L2 round trip]
//   P3  (workers) INPUT-centric resampling: source j knows its cdf value c_j, hence -- in closed form --
//       the number F(c_j) of output slots whose position (i + U_i)/N is <= c_j; it owns the output
//       slots [F(c_{j-1}), F(c_j)).  No search.  Warp boundaries are F of values both neighbours
//       read from the same shared-memory (CTA boundaries: the same L2) words, so every slot is
//       produced exactly once without any cross-warp prefix
//   P4  WARP-PRIVATE expansion: every source marks the first of its slots in the warp's own
//       head array, a running maximum over the slots tells every slot its source, the chosen x are
//       staged per warp and leave the SM as coalesced 16-byte LL stores into x_new.  No block barrier
//   B2  every thread polls its own elements of x_new: value and epoch tag travel in one 8-byte word
//       (LL protocol), so there is no fence, no flag, no barrier; a warp whose particles have
//       arrived starts the next observation at once                          [< 1 L2 round trip]
// Only x_new (one write + one read per particle, L2 resident) and the tiny records leave the SM.
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
constexpr int FAST_BAR_A = 1;        // named barriers: workers -> service, service -> workers
constexpr int FAST_BAR_B = 2;
// output slots per lane of one expansion pass: 25 % beyond the lane's sources, rounded up to whole 16-byte accesses
__host__ __device__ constexpr int fast_spt(int ppt) { return (ppt * 5 / 4 + 3) & ~3; }
// stride (elements) of a lane's particles in the warp's staging array: + 4 keeps the 16-byte accesses conflict-free
__host__ __device__ constexpr int fast_xs(int ppt) { return ppt % 8 == 4 ? ppt : ppt + 4; }

// LL ("low latency") words: 32 data bits + 32-bit epoch tag in one 8-byte unit, two units per
// 16-byte access.  A reader that sees the expected tag also sees the data: no fence, no separate
// flag, no dependent second load.  The accesses are relaxed at GPU scope (all that an exchange between
// CTAs of one GPU needs).
// The group exchange: one record per CTA and observation, polled by every CTA of the group.  A record holds what the
// resampling decision and the cdf need (max, sum, sum of squares of the CTA's weights) in CU 16-byte units; the state
// sums -- outputs only CTA 0 writes -- travel in a second array that CTA 0 reads after it has released its workers.
// (scripts/probes/exchange_probe.cu: one store -> remote load hop through the B200's L2 costs about 1060 cycles, an
// all-to-all of 148 CTAs about 4200; a private inbox per reader, G x G stores, was slower: 6350.)
template <bool F32> struct FastRecLayout {
  static constexpr int CU = F32 ? 2 : 3;      // units per record: f32 (s), (m, q); f64 m, s, q
  static constexpr int CUS = F32 ? 2 : 4;     // stride in units
};
__host__ __device__ inline size_t fast_rec_units(int ngroups, int G, int cus) { return (size_t)ngroups * 2 * G * cus; }
__host__ __device__ inline size_t fast_aux_units(int ngroups, int G) { return (size_t)ngroups * 2 * G * 2; }

// launch geometry, shared by fast_launch() (bssm_fast.cu) and the CPU logic tests
struct FastGeom {
  int nb_max, nw, threads, ch, ucap, uw;
  size_t smem;
};
template <typename Real, int PPT>
inline FastGeom fast_geometry(int N, int G, int uw_req) {
  FastGeom g;
  g.nb_max = (N + G - 1) / G;
  g.nb_max = (g.nb_max + PPT - 1) / PPT * PPT;
  g.nw = (g.nb_max + 32 * PPT - 1) / (32 * PPT);
  if (g.nw < 1) g.nw = 1;
  g.threads = (g.nw + 1) * 32;
  g.ch = 32 * fast_spt(PPT);
  g.uw = uw_req >= 0 ? uw_req : (g.nb_max / 7 < 64 ? 64 : (g.nb_max / 7 > 2048 ? 2048 : g.nb_max / 7));
  g.uw = (g.uw + 3) & ~3;
  g.ucap = (g.nb_max + 2 * g.uw + 3) & ~3;
  // parity precision with big slices: the per-warp staging areas fill the shared memory; no CTA-wide window of uniforms,
  // every warp stages its own (the window is an optimisation only)
  if (sizeof(Real) == 8 && g.nw > 14 && uw_req < 0) { g.uw = 0; g.ucap = 0; }
  g.smem = (size_t)(160 + 34 + 32) * sizeof(double) + 16 + (size_t)((5 * G + 1) & ~1) * sizeof(double) + (size_t)2 * g.ucap * sizeof(unsigned int) +
           (size_t)g.nw * ((size_t)32 * fast_xs(PPT) * sizeof(Real) + (size_t)g.ch * sizeof(unsigned int) + (size_t)g.ch * sizeof(Real));
  return g;
}

struct FastParams {
  FilterDev f;
  int G, ngroups;
  int resample_fn;
  uint4* rec;       // [ngroups][2][G][CUS] LL units: the CTAs' records
  uint4* aux;       // [ngroups][2][G][2] LL units: the state sums, read by CTA 0
  void* xnew;       // [ngroups][G * nb_max] LL elements: uint2 (f32) / uint4 (f64)
  int nb_max;       // slice stride (multiple of PPT)
  int ucap, uw;     // staged window of stratified uniforms: slots [base - uw, base - uw + ucap) of the CTA with first particle `base`
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
// named barriers (PTX bar.sync / bar.arrive with a thread count): the documented producer / consumer pairing --
// memory accesses before the arrive are performed before the matching sync returns
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive_named(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
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

template <typename Real> __device__ __forceinline__ Real fast_warp_sum(Real v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename Model, typename Real, int PPT, int NWMAX>
__global__ void __launch_bounds__((NWMAX + 1) * 32, 1) k_fast_bpf(FastParams P) {
  static_assert(Model::D == 1 && Model::NZ_TRANS == 1 && Model::NU_TRANS == 0 && Model::NZ_INIT == 1 && Model::NU_INIT == 0,
                "persistent kernel: 1-D models with one normal per transition");
  static_assert(PPT % 4 == 0, "one Philox call serves 4 particles");
  constexpr bool F32 = sizeof(Real) == 4;
  constexpr int SPT = fast_spt(PPT);                 // output slots per lane in one expansion pass
  constexpr int CH = 32 * SPT;                       // ... per warp
  constexpr int XS = fast_xs(PPT);                   // stride of a lane's particles in the warp's staging array
  constexpr int VR = 16 / (int)sizeof(Real);         // Reals per 16-byte access
  static_assert(SPT % 4 == 0 && XS % 8 == 4 && 32 * XS <= 1024, "16-byte accesses to the head / staging arrays; 10-bit source index");
  typedef FastRecLayout<F32> RL;
#ifndef BSSM_EMU
  extern __shared__ __align__(16) unsigned char smem_raw[];
#else
  unsigned char* const smem_raw = emu_dynamic_smem();
#endif
  const FilterDev& f = P.f;
  const int G = P.G;
  const int group = blockIdx.x / G, b = blockIdx.x % G;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int NW = ((int)blockDim.x >> 5) - 1;         // worker warps; warp NW is the service warp
  const int NBAR = (NW + 1) * 32;
  // shared memory carve-up
  double* s_wrec = (double*)smem_raw;                // [5][32] one record per worker warp: m, s, q, sx, pending
  double* s_T0 = s_wrec + 160;                       // [NW + 1] slot position of the first particle of every warp (and of the next CTA)
  double* s_sl = s_T0 + 34;                          // [NW] output slots per unit of the warp's e
  int* s_flag = (int*)(s_sl + 32);                   // dead | resample << 1
  double* s_tab = (double*)(s_flag + 4);             // [5][G] (service warp) the group's records
  unsigned int* s_u = (unsigned int*)(s_tab + ((5 * G + 1) & ~1));   // [2][ucap] staged stratified uniforms (raw words), by observation parity
  unsigned char* s_warp = (unsigned char*)(s_u + 2 * P.ucap);
  constexpr size_t WARP_BYTES = (size_t)32 * XS * sizeof(Real) + (size_t)CH * sizeof(unsigned int) + (size_t)CH * sizeof(Real);
  Real* s_xs = (Real*)(s_warp + (size_t)(wid < NW ? wid : 0) * WARP_BYTES);   // [32][XS] this warp's particles
  unsigned int* s_hd = (unsigned int*)(s_xs + 32 * XS);                         // [CH] expansion: epoch << 10 | index into s_xs, at the first slot of a source
  Real* s_out = (Real*)(s_hd + CH);                                             // [CH] staging of the chosen x (also: the warp's own window of uniforms)

  uint4* const rec = P.rec + (size_t)group * 2 * G * RL::CUS;
  uint4* const aux = P.aux + (size_t)group * 2 * G * 2;
  typedef typename std::conditional<F32, uint2, uint4>::type XEl;   // LL element of x_new
  XEl* xnew = (XEl*)P.xnew + (size_t)group * G * P.nb_max;
  const double INF = __longlong_as_double(0x7FF0000000000000LL), NINF = -INF;
#ifdef BSSM_FAST_TIMING_BUILD
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define FAST_TICK(ph) do { if (P.timing && lane == 0) { long long t_ = clock64(); tacc[ph] += t_ - tprev; tprev = t_; } } while (0)
  // absolute times (ns, %globaltimer) in observations [500, 532): the service warp's exchange and worker warp 5's phases
#define FAST_TRACE(slot, obs_) do { if (P.timing && lane == 0 && (wid == NW || wid == (NW > 5 ? 5 : 0)) && (obs_) >= 500 && (obs_) < 532) { unsigned long long g_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_)); \
    P.timing[(size_t)gridDim.x * 128 + ((size_t)blockIdx.x * 32 + ((obs_) - 500)) * 12 + (slot)] = (long long)g_; } } while (0)
#else
#define FAST_TICK(ph) do { } while (0)
#define FAST_TRACE(slot, obs_) do { } while (0)
#endif

  if (wid == NW) {
    // =========================== service warp ===========================
    unsigned int ep1 = 0;          // record epoch: the same sequence in every CTA of the group
    const int R = (G + 31) >> 5;   // records per lane
    for (int c = group; c < f.C; c += P.ngroups) {
      if (!f.alive[c]) continue;
      const int n = filt_n(f, c);
      int nb = (n + G - 1) / G;
      nb = (nb + PPT - 1) / PPT * PPT;
      const int base = b * nb;
      const int T1 = f.T + 1;
      const int ralg = f.ralg;
      double thr = f.threshold;
      if (thr < 0) thr = (ralg == 0) ? INF : (ralg == 1 ? (double)n : (double)n / 2.0);
      const double log_n = log((double)n);
      double loglike = 0.0;
      int n_resampled = 0, pending_obs = -1;
      // phase: 0 = t = 0 state estimate, 1 = observation `obs`, 2 = final flush.  Returns dead | resample << 1
      auto exchange = [&](int phase, int obs) -> int {
        FAST_TICK(0);
        bar_sync_named(FAST_BAR_A, NBAR);
        const bool act = lane < NW;
        const double mw = act ? *(volatile double*)&s_wrec[lane] : NINF;
        FAST_TICK(1);   // wait for the workers' records
        FAST_TRACE(0, obs);
        // ---- the CTA's record: rescale the warp records to the CTA maximum, exclusive prefix over the warps.  First what the
        //      decision needs (max, sum, sum of squares); the state sums follow after the publication ----
        double mb;
        if (F32) {   // the warp maxima are floats: one 32-bit shuffle per step
          float t = (float)mw;
#pragma unroll
          for (int o = 16; o; o >>= 1) { const float u = __shfl_xor_sync(0xffffffffu, t, o); t = u > t ? u : t; }
          mb = (double)t;
        } else mb = warp_max_d(mw);
        double scw = 0.0;
        if (!(mw == NINF || mb == NINF)) scw = F32 ? (double)__expf((float)(mw - mb)) : exp(mw - mb);
        const double sbw = act ? s_wrec[32 + lane] * scw : 0.0;
        const double qbw = act ? s_wrec[64 + lane] * scw * scw : 0.0;
        const double inc_w = warp_incl_scan_d(sbw, lane);
        const double q_b = F32 ? (double)fast_warp_sum<float>((float)qbw) : warp_sum_d(qbw);
        const double wex = inc_w - sbw;
        const double s_b = __shfl_sync(0xffffffffu, inc_w, 31);
        ep1++;
        const int par = (int)(ep1 & 1u);
        if (G > 1) {
          uint4* dst = rec + (size_t)(par * G + b) * RL::CUS;
          if (F32) {
            if (lane == 0) ll_put_double(dst, s_b, ep1);
            else if (lane == 1) ll_store_v4(dst + 1, __float_as_uint((float)mb), ep1, __float_as_uint((float)q_b), ep1);
          } else if (lane < 3) ll_put_double(dst + lane, lane == 0 ? mb : (lane == 1 ? s_b : q_b), ep1);
        }
        FAST_TICK(2);   // CTA record + publish
        FAST_TRACE(1, obs);
        // the state sums: to CTA 0, which reads them after it has released its workers
        const double x_b = warp_sum_d(act ? s_wrec[96 + lane] * scw : 0.0);
        const double p_b = warp_sum_d(act ? s_wrec[128 + lane] : 0.0);
        if (G > 1) {
          uint4* dst = aux + (size_t)(par * G + b) * 2;
          if (F32) { if (lane == 0) ll_store_v4(dst, __float_as_uint((float)x_b), ep1, __float_as_uint((float)p_b), ep1); }
          else if (lane < 2) ll_put_double(dst + lane, lane == 0 ? x_b : p_b, ep1);
          // ---- poll the G records: every load of a round is in flight at once ----
          const uint4* src = rec + (size_t)par * G * RL::CUS;
          const int nunits = G * RL::CUS;
          constexpr int UB = 10;   // units per lane and round
          for (int u0 = 0; u0 < nunits; u0 += 32 * UB) {
            uint4 v[UB];
            bool ok;
            do {
              ok = true;
#pragma unroll
              for (int i = 0; i < UB; i++) {
                const int u = u0 + 32 * i + lane;
                if (u < nunits && (u % RL::CUS) < RL::CU) { v[i] = ll_load_v4(src + u); ok = ok && v[i].y == ep1 && v[i].w == ep1; }
              }
            } while (!ok);
#pragma unroll
            for (int i = 0; i < UB; i++) {
              const int u = u0 + 32 * i + lane;
              if (u < nunits) {
                const int j = u / RL::CUS, h = u % RL::CUS;
                if (F32) {
                  if (h == 0) s_tab[G + j] = ll_get_double(v[i]);
                  else { s_tab[j] = (double)__uint_as_float(v[i].x); s_tab[2 * G + j] = (double)__uint_as_float(v[i].z); }
                } else if (h < 3) s_tab[h * G + j] = ll_get_double(v[i]);
              }
            }
          }
        } else if (lane == 0) {
          // a group of one: the CTA's record is the group's (rounded like a published one, so that G never changes a result by more than summation order)
          if (F32) { s_tab[0] = (double)(float)mb; s_tab[1] = s_b; s_tab[2] = (double)(float)q_b; }
          else { s_tab[0] = mb; s_tab[1] = s_b; s_tab[2] = q_b; }
        }
        __syncwarp();
        FAST_TICK(3);   // poll
        FAST_TRACE(2, obs);
        // ---- global max / sums / this CTA's cdf interval: R consecutive records per lane, one warp scan; the same
        //      expressions on the same table in every CTA, so neighbouring CTAs agree on their common boundary bit for bit ----
        double M = NINF;
        for (int j = lane; j < G; j += 32) M = s_tab[j] > M ? s_tab[j] : M;
        if (F32) {
          float t = (float)M;
#pragma unroll
          for (int o = 16; o; o >>= 1) { const float u = __shfl_xor_sync(0xffffffffu, t, o); t = u > t ? u : t; }
          M = (double)t;
        } else M = warp_max_d(M);
        const int j0 = lane * R;
        double loc_s = 0.0, loc_q = 0.0, my_lo = 0.0, my_hi = 0.0, my_g = 0.0;
        for (int r = 0; r < R; r++) {
          const int j = j0 + r;
          if (j < G) {
            const double mj = s_tab[j];
            double sc = 0.0;
            if (!(mj == NINF || M == NINF)) sc = F32 ? (double)__expf((float)(mj - M)) : exp(mj - M);
            s_tab[3 * G + j] = sc;                    // kept for the state sums
            loc_s += s_tab[G + j] * sc;
            if (j == b - 1) my_lo = loc_s;            // lane-local inclusive values of records b-1 and b
            if (j == b) { my_hi = loc_s; my_g = sc; }
            loc_q += s_tab[2 * G + j] * sc * sc;
          }
        }
        const double inc = warp_incl_scan_d(loc_s, lane);
        const double Q = warp_sum_d(loc_q);
        const double off = inc - loc_s;               // everything before this lane's first record
        const double S = __shfl_sync(0xffffffffu, inc, 31);
        const double A_hi = __shfl_sync(0xffffffffu, off + my_hi, b / R);
        const double A_lo = b == 0 ? 0.0 : __shfl_sync(0xffffffffu, off + my_lo, (b - 1) / R);
        const double g_b = __shfl_sync(0xffffffffu, my_g, b / R);
        int dead = 0, resample = 0;
        bool bad = false, empty = false;
        if (phase == 1) {
          bad = (S != S) || (M != M);
          empty = M < -1e8;
          dead = (bad || empty) ? 1 : 0;
          // ess < thr  <=>  S^2 < thr * Q  (no division on the critical path)
          resample = dead ? 0 : ((ralg == 0) ? 0 : (ralg == 1 ? 1 : (S * S < thr * Q)));
          if (resample) {
            // slot position of the first particle of every worker warp, and of the first particle after this CTA;
            // a warp that starts at or beyond particle n sits at the end of the slots
            const double nS = (double)n / S;
            const long long lstart = (long long)lane * 32 * PPT < (long long)nb ? (long long)lane * 32 * PPT : (long long)nb;   // first particle of warp `lane` within the slice
            double t0 = lstart >= nb ? A_hi * nS : (A_lo + wex * g_b) * nS;    // a warp beyond the slice starts where the next CTA does
            if ((long long)base + lstart >= (long long)n) t0 = 2.0 * (double)n;
            if (lane <= NW) s_T0[lane] = t0;
            if (lane < NW) s_sl[lane] = scw * g_b * nS;
          }
        }
        if (lane == 0) s_flag[0] = dead | (resample << 1);
        bar_arrive_named(FAST_BAR_B, NBAR);
        FAST_TICK(4);   // merge
        FAST_TRACE(3, obs);
        if (b == 0) {
          // the outputs of this observation, off everybody's critical path: the state sums of the group (CTA 0 only)
          double SX, PEND;
          if (G > 1) {
            double lx = 0.0, lp = 0.0;
            const uint4* src = aux + (size_t)par * G * 2;
            for (int j = lane; j < G; j += 32) {
              uint4 v0, v1 = make_uint4(0u, 0u, 0u, 0u);
              bool ok;
              do {
                v0 = ll_load_v4(src + 2 * j);
                ok = v0.y == ep1 && v0.w == ep1;
                if (!F32) { v1 = ll_load_v4(src + 2 * j + 1); ok = ok && v1.y == ep1 && v1.w == ep1; }
              } while (!ok);
              const double sxj = F32 ? (double)__uint_as_float(v0.x) : ll_get_double(v0);
              const double pj = F32 ? (double)__uint_as_float(v0.z) : ll_get_double(v1);
              lx += sxj * s_tab[3 * G + j]; lp += pj;
            }
            SX = warp_sum_d(lx); PEND = warp_sum_d(lp);
          } else {
            SX = F32 ? (double)(float)x_b : x_b; PEND = F32 ? (double)(float)p_b : p_b;
          }
          if (lane == 0) {
            if (phase == 0) {
              f.ess[(size_t)c * T1] = (double)n;
              f.state_est[(size_t)c * T1] = SX / (double)n;
            } else {
              if (pending_obs >= 0) f.state_est[(size_t)c * T1 + pending_obs + 1] = PEND / (double)n;
              if (phase == 1) {
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
              } else {
                f.loglike[c] = loglike; f.n_resampled[c] = n_resampled;
              }
            }
          }
        }
        pending_obs = resample ? obs : -1;
        n_resampled += resample;
        return dead | (resample << 1);
      };
      exchange(0, -1);
      for (int obs = 0; obs < f.T; obs++) {
        if (exchange(1, obs) & 1) break;
      }
      exchange(2, -1);
    }
  } else {
    // =========================== worker warps ===========================
    const int wt = tid;            // worker thread index
    const int NWT = NW * 32;
    unsigned int ep2 = 0;          // x_new epoch: the same sequence in every CTA of the group
    unsigned int hep = 0;          // head-array epoch of this warp (never reset: stale heads always compare low)
    for (int i = lane; i < CH; i += 32) s_hd[i] = 0u;
    __syncwarp();
    for (int c = group; c < f.C; c += P.ngroups) {
      if (!f.alive[c]) continue;
      const int n = filt_n(f, c);
      int nb = (n + G - 1) / G;
      nb = (nb + PPT - 1) / PPT * PPT;
      const int base = b * nb;                                  // first global particle of this CTA
      const int ibase = base + wt * PPT;                        // first global particle of this thread
      const int n_own = max(0, min(PPT, min(n - ibase, nb - wt * PPT)));  // owned particles of this thread
      Real par[Model::NPAR];
      Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
      const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
      const int ralg = f.ralg;
      const int u_base = max(0, (base - P.uw) & ~3);            // first slot of the CTA's staged window of uniforms

      // ---- init (R/particle_filter_core.R:76-116) ----
      Real x[PPT];
      Real px = 0;   // sum of this thread's particles after the last resampling (state estimate, travels in the next record)
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
          px += x[4 * h + k];
        }
      }
      // t = 0 state estimate: the sum of the initial particles travels in the record's sx
      {
        const double v = warp_sum_d((double)px);
        if (lane == 0) { s_wrec[wid] = 0.0; s_wrec[32 + wid] = 0.0; s_wrec[64 + wid] = 0.0; s_wrec[96 + wid] = v; s_wrec[128 + wid] = 0.0; }
        bar_sync_named(FAST_BAR_A, NBAR);
        bar_sync_named(FAST_BAR_B, NBAR);
        px = 0;
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

        FAST_TICK(0);   // reload of x_new (resample steps)
        FAST_TRACE(8, obs);
        // ---- P1: propagate + log-weight ----
        for (int tnow = prev_t + 1; tnow <= ot; tnow++) {
          if (zpre_t != tnow - 1) gen_normals(tnow - 1);
#pragma unroll
          for (int k = 0; k < PPT; k++) {
            Real zi[1] = {zpre[k]};
            Model::template transition<Real>(&x[k], par, tnow, zi, nullptr);
          }
        }
        Real e[PPT];   // first the log-weights, then exp(lw - warp max)
        Real mloc = Math<Real>::ninf();
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          e[k] = Model::template loglik<Real>(yv, &x[k], par, ot);
          if (k >= n_own) e[k] = Math<Real>::ninf();
          mloc = e[k] > mloc ? e[k] : mloc;
        }
        // warp max (a NaN log-weight is not an ordered maximum: it reaches the sums through exp below)
        Real mw = mloc;
#pragma unroll
        for (int o = 16; o; o >>= 1) { Real t = __shfl_xor_sync(0xffffffffu, mw, o); mw = t > mw ? t : mw; }
        double exu;            // warp-local exclusive prefix of this thread (unnormalised, relative to the warp max)
        {
          Real fs = 0, fq = 0, fx = 0;
          const Real mr = (mw == Math<Real>::ninf()) ? (Real)0 : mw;   // exp(-inf - 0) = 0: no per-particle guard
#pragma unroll
          for (int k = 0; k < PPT; k++) {
            Real ek = Math<Real>::exp_(e[k] - mr);
            e[k] = ek;
            fs += ek; fq += ek * ek; fx += ek * x[k];
          }
          const double run = (double)fs;
          const double inc = warp_incl_scan_d(run, lane);
          exu = inc - run;
          const double ws = __shfl_sync(0xffffffffu, inc, 31);
          const Real tq = fast_warp_sum<Real>(fq), tx = fast_warp_sum<Real>(fx), tp = fast_warp_sum<Real>(px);
          if (lane == 0) { s_wrec[wid] = (double)mw; s_wrec[32 + wid] = ws; s_wrec[64 + wid] = (double)tq; s_wrec[96 + wid] = (double)tx; s_wrec[128 + wid] = (double)tp; }
        }
        FAST_TICK(1);   // P1 (propagate, weights, warp reductions)
        FAST_TRACE(9, obs);
        // all records are in: the service warp runs the exchange, the workers the part of the resampling and of the next
        // observation that does not depend on it.  A full barrier rather than an arrive: the next normals of a fast warp must
        // not take issue slots from a slower warp's weights, which are on the critical path
        bar_sync_named(FAST_BAR_A, NBAR);
        FAST_TICK(2);   // wait for the other workers
        // ---- work that does not depend on the exchange ----
        if (obs + 1 < f.T) gen_normals(ot);
        unsigned int w_sys = 0u;
        unsigned int* const su = s_u + (obs & 1) * P.ucap;
        if (ralg != 0) {
          if (P.resample_fn == 1) {
            uint4x q0 = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, 0u);
            w_sys = q0.w[0];
          } else {
            // stage the Philox words of the slots this CTA will probably serve: one call per 4 slots
            const int q_end = min((n + 3) >> 2, (u_base + P.ucap) >> 2);
            for (int qd = (u_base >> 2) + wt; qd < q_end; qd += NWT) {
              uint4x uq = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, (unsigned int)qd);
              *(uint4*)&su[4 * qd - u_base] = make_uint4(uq.w[0], uq.w[1], uq.w[2], uq.w[3]);
            }
          }
        }
        FAST_TICK(3);   // next-step normals + uniforms
        bar_sync_named(FAST_BAR_B, NBAR);
        const int fl = *(volatile int*)&s_flag[0];
        FAST_TICK(4);   // wait for the exchange
        FAST_TRACE(4, obs);
        px = 0;
        if (fl & 1) break;
        if (!(fl & 2)) continue;

        // ---- P3: closed-form offspring ranges ----
        SlotCounter sc;
        sc.key = key; sc.obs = (unsigned int)obs; sc.fn = P.resample_fn; sc.n = n; sc.w_sys = w_sys;
        sc.s_u = su; sc.u_base = u_base; sc.u_cap = P.ucap;
        // output range of this warp: F at the warp's first particle and at the next warp's (both neighbours evaluate the same words)
        int o_start, o_end;
        {
          const double tb = s_T0[wid + (lane & 1)];
          const int o = sc.count_slots(tb);
          o_start = __shfl_sync(0xffffffffu, o, 0);
          o_end = __shfl_sync(0xffffffffu, o, 1);
          if (o_end < o_start) o_end = o_start;   // cannot happen with a monotone table; keeps the ranges sane if it ever did
        }
        if (P.resample_fn != 1 && (o_start - 1 < u_base || o_end + 1 > u_base + P.ucap) && o_end > o_start) {
          // the warp's slots fell outside the CTA's window: stage the warp's own window (what does not fit is recomputed per slot)
          const int wb = max(0, (o_start - 1) & ~3);
          unsigned int* const wu = (unsigned int*)s_out;
          const int q_end = min((n + 3) >> 2, (wb + CH) >> 2);
          for (int qd = (wb >> 2) + lane; qd < q_end; qd += 32) {
            uint4x uq = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, (unsigned int)qd);
            *(uint4*)&wu[4 * qd - wb] = make_uint4(uq.w[0], uq.w[1], uq.w[2], uq.w[3]);
          }
          __syncwarp();
          sc.s_u = wu; sc.u_base = wb; sc.u_cap = CH;
        }
        // F of this thread's sources (monotone by a running max; clamped into [o_start, o_end])
        int F[PPT];
        {
          const double sl = s_sl[wid];
          const double T0 = s_T0[wid] + exu * sl;      // slot position just before this thread's first particle
          int fmax = o_start;
          if (F32) {
            // fp64 origin per thread, fp32 increments: t_k = T0 + (sum of e up to k) * sl
            const double T0c = T0 < (double)n ? T0 : (double)n;
            const int I0 = (int)T0c;
            const float f0 = (float)(T0c - (double)I0);
            const float wsn = (float)sl;
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
              if (ibase + k == n - 1 || (lane == 31 && k == PPT - 1)) v = o_end;   // the warp's last particle takes what is left
              v = min(max(v, o_start), o_end);
              if (k >= n_own) v = o_start;
              fmax = max(fmax, v);
              F[k] = fmax;
            }
          } else {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < PPT; k++) {
              acc += (double)e[k];
              int v = o_start;
              if (k < n_own) {
                v = sc.count_slots(T0 + acc * sl);
                if (ibase + k == n - 1 || (lane == 31 && k == PPT - 1)) v = o_end;
                v = min(max(v, o_start), o_end);
              }
              fmax = max(fmax, v);
              F[k] = fmax;
            }
          }
        }
        // exclusive prefix-max of the per-lane last F over the warp
        int prevF;
        {
          int inc = F[PPT - 1];
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, t); }
          prevF = __shfl_up_sync(0xffffffffu, inc, 1);
          if (lane == 0) prevF = o_start;
#pragma unroll
          for (int k = 0; k < PPT; k++) F[k] = max(F[k], prevF);
        }
        FAST_TICK(5);   // offspring ranges
        FAST_TRACE(6, obs);
        // ---- P4: warp-private expansion, output-centric (passes of CH slots).  Every source with offspring in the pass marks
        //      the first of its slots with (epoch, index of its x in s_xs); a running maximum over the slots -- source indices grow
        //      with the slot, older epochs compare low -- tells every slot its source: O(1) per source and per slot, no loop over the
        //      offspring of a source, no special case for heavy sources, nothing to clear ----
        {
          hep++;
          const unsigned int hkey = hep << 10;
#pragma unroll
          for (int h4 = 0; h4 < PPT / VR; h4++) {
            if (F32) *(float4*)&s_xs[lane * XS + 4 * h4] = make_float4((float)x[4 * h4], (float)x[4 * h4 + 1], (float)x[4 * h4 + 2], (float)x[4 * h4 + 3]);
            else *(double2*)&s_xs[lane * XS + 2 * h4] = make_double2((double)x[2 * h4], (double)x[2 * h4 + 1]);
          }
          const unsigned int tag = ep2 + 1;
          unsigned int pass_carry = 0u;
          for (int c0 = o_start & ~3; c0 < o_end; c0 += CH) {
            const int c1 = min(o_end, c0 + CH);
            int lo_k = prevF;
#pragma unroll
            for (int k = 0; k < PPT; k++) {
              const int hi_k = F[k];
              const int a = max(lo_k, c0);
              if (min(hi_k, c1) > a) s_hd[a - c0] = hkey | (unsigned int)(lane * XS + k);
              lo_k = hi_k;
            }
            __syncwarp();
            unsigned int hd[SPT];
#pragma unroll
            for (int i = 0; i < SPT; i += 4) { const uint4 v4 = *(const uint4*)&s_hd[lane * SPT + i]; hd[i] = v4.x; hd[i + 1] = v4.y; hd[i + 2] = v4.z; hd[i + 3] = v4.w; }
#pragma unroll
            for (int i = 1; i < SPT; i++) hd[i] = max(hd[i], hd[i - 1]);
            unsigned int inc = hd[SPT - 1];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, t); }
            unsigned int carry = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) carry = 0u;
            carry = max(carry, pass_carry);
            pass_carry = max(pass_carry, __shfl_sync(0xffffffffu, inc, 31));
            Real val[SPT];
#pragma unroll
            for (int i = 0; i < SPT; i++) val[i] = s_xs[max(hd[i], carry) & 1023u];
#pragma unroll
            for (int i = 0; i < SPT; i += VR) {
              if (F32) *(float4*)&s_out[lane * SPT + i] = make_float4((float)val[i], (float)val[i + 1], (float)val[i + 2], (float)val[i + 3]);
              else *(double2*)&s_out[lane * SPT + i] = make_double2((double)val[i], (double)val[i + 1]);
            }
            __syncwarp();
            // copy out as LL elements (value + epoch tag): 16-byte stores, 8-byte at the ragged ends
            const int first = max(c0, o_start), last = c1;   // slots [first, last) are valid in this pass
            if (F32) {
#pragma unroll
              for (int it = 0; it < SPT / 2; it++) {
                const int o = c0 + 2 * (it * 32 + lane);
                const float2 v = *(const float2*)&s_out[o - c0];
                if (o >= first && o + 1 < last) ll_store_v4(&xnew[o], __float_as_uint(v.x), tag, __float_as_uint(v.y), tag);
                else {
                  if (o >= first && o < last) ll_store_v2(&xnew[o], __float_as_uint(v.x), tag);
                  if (o + 1 >= first && o + 1 < last) ll_store_v2(&xnew[o + 1], __float_as_uint(v.y), tag);
                }
              }
            } else {
#pragma unroll
              for (int it = 0; it < SPT; it++) {
                const int o = c0 + it * 32 + lane;
                if (o >= first && o < last) ll_put_double((uint4*)&xnew[o], (double)s_out[o - c0], tag);
              }
            }
            __syncwarp();
          }
        }
        ep2++;
        FAST_TICK(6);   // expansion + LL copy-out
        FAST_TRACE(7, obs);
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
#pragma unroll
          for (int k = 0; k < PPT; k++) px += x[k];   // state estimate after resampling: travels in the next record
        }
      }  // obs
      // flush: the state estimate of a final resampling step still travels in the records
      {
        const double v = warp_sum_d((double)px);
        if (lane == 0) { s_wrec[wid] = 0.0; s_wrec[32 + wid] = 0.0; s_wrec[64 + wid] = 0.0; s_wrec[96 + wid] = 0.0; s_wrec[128 + wid] = v; }
        bar_sync_named(FAST_BAR_A, NBAR);
        bar_sync_named(FAST_BAR_B, NBAR);
      }
    }    // filters
  }
#ifdef BSSM_FAST_TIMING_BUILD
  if (P.timing && lane == 0 && wid < 16) for (int i = 0; i < 8; i++) P.timing[((size_t)blockIdx.x * 16 + wid) * 8 + i] = tacc[i];
#endif
#undef FAST_TICK
#undef FAST_TRACE
}

}  // namespace bssm
