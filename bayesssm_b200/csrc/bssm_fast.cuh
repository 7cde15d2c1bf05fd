// bssm_fast.cuh -- persistent bootstrap-filter kernel: the whole T loop of
// .particle_filter_core (R/particle_filter_core.R:123-246) for algorithm "BPF" in ONE launch.
//
// A filter is run by a GROUP of G co-resident CTAs (cooperative launch); CTA b owns the contiguous slice
// [b*nb, (b+1)*nb) of the particles and keeps it in REGISTERS for all T steps (PPT particles per thread).
// Per observation:
//   P1  propagate (normals pre-generated while the previous record travelled; one Philox call per 4
//       particles) + log-weight; e = exp(lw - ref) against a reference every CTA knows WITHOUT
//       communication -- the model's upper bound of the log-likelihood -- so the weights of all CTAs
//       are on one scale from the start: block sums of e, e^2, e*x, block max of lw and the block-local
//       inclusive scan of e                                   [registers + shuffles, 1 block barrier]
//   B1  every CTA publishes (max_b, sums) as an epoch-stamped LL record in L2, generates the NEXT
//       step's normals while the record travels, then polls the G records (no atomics, no fences): the
//       polling warps scan the records of 32 CTAs each, and after one block barrier every warp adds up at
//       most 8 group totals to get -- bit-identically in every warp of every CTA -- the global sum / ESS /
//       resampling decision, the cdf interval of its CTA and the slot positions where its own particles
//       begin and end.  (If the best particle lies more than e^-60 below the bound -- an outlying
//       observation -- the step is redone against the true maximum, which the records carry.)
//                                                                [1 L2 all-to-all, 1 block barrier]
//   P3  INPUT-centric resampling: source j knows its cdf value c_j, hence -- in closed form -- the
//       number F(c_j) of output slots whose position (i + U_i)/N is <= c_j; it owns the output
//       slots [F(c_{j-1}), F(c_j)).  No search.  Warp boundaries are F of values both neighbouring
//       warps (CTAs) compute with the same expression from the same shared-memory (L2) words, so
//       every slot is produced exactly once without any cross-warp prefix or barrier.  Each warp
//       stages the stratified uniforms of exactly its own output range (one Philox call per 4 slots)
//   P4  WARP-PRIVATE expansion: every source marks the first of its slots in the warp's own
//       head array, a running maximum over the slots tells every slot its source, the chosen x are
//       staged per warp and leave the SM as coalesced 16-byte LL stores into x_new.  No block barrier
//   B2  every thread polls its own elements of x_new: value and epoch tag travel in one 8-byte word
//       (LL protocol), so there is no fence, no flag, no barrier; a warp whose particles have
//       arrived starts the next observation at once                          [1 L2 hop]
// Only x_new (one write + one read per particle, L2 resident) and the tiny records leave the SM.
// Same Philox keying and tie rule (first j with cdf[j] >= pos, clamp) as the general engine, so
// results do not depend on G or the launch geometry beyond floating-point summation order.
// In the throughput precision (Real = float) the within-thread part of the cdf and the slot
// arithmetic run in fp32 relative to an fp64 per-thread origin.
// What the B200 charges for the exchanges (scripts/probes/exchange_probe.cu): one store -> remote poll hop
// through L2 about 1060 cycles, the all-to-all of 148 CTAs about 4200 per round.
#pragma once
#include "bssm_common.cuh"
#include "bssm_filter.cuh"
#include "bssm_models.cuh"
#include "bssm_slots.cuh"

#include <type_traits>

namespace bssm {

constexpr int FAST_MAX_NB = 7168;    // particles per CTA
constexpr int FAST_MAX_G = 256;      // CTAs per group
// output slots per lane of one expansion pass: 25 % beyond the lane's sources, rounded up to whole 16-byte accesses
__host__ __device__ constexpr int fast_spt(int ppt) { return (ppt * 5 / 4 + 3) & ~3; }
// stride (elements) of a lane's particles in the warp's staging array: + 4 keeps the 16-byte accesses conflict-free
__host__ __device__ constexpr int fast_xs(int ppt) { return ppt % 8 == 4 ? ppt : ppt + 4; }

// LL ("low latency") words: 32 data bits + 32-bit epoch tag in one 8-byte unit, two units per
// 16-byte access.  A reader that sees the expected tag also sees the data: no fence, no separate
// flag, no dependent second load.  The accesses are relaxed at GPU scope (all that an exchange between
// CTAs of one GPU needs).
// The group exchange: one record per CTA and observation, polled by every CTA of the group: NU 16-byte units.
// f32: (sum), (max, sum of squares), (state sum, pending state sum); f64: five doubles
template <bool F32> struct FastRecLayout {
  static constexpr int NU = F32 ? 3 : 5;
  static constexpr int NUS = F32 ? 4 : 8;     // stride in units (64 / 128 bytes)
};
__host__ __device__ inline size_t fast_rec_units(int ngroups, int G, int nus) { return (size_t)ngroups * 2 * G * nus; }

// launch geometry, shared by fast_launch() (bssm_fast.cu) and the CPU logic tests
struct FastGeom {
  int nb_max, nw, threads, ch, ucap, uw, xstride;
  size_t smem;
};
template <typename Real, int PPT>
inline FastGeom fast_geometry(int N, int G, int uw_req /* < 0: default */) {
  FastGeom g;
  g.nb_max = (N + G - 1) / G;
  g.nb_max = (g.nb_max + PPT - 1) / PPT * PPT;
  g.nw = (g.nb_max + 32 * PPT - 1) / (32 * PPT);
  if (g.nw < 1) g.nw = 1;
  g.threads = g.nw * 32;
  g.ch = 32 * fast_spt(PPT);
  // uw_req < 0: no CTA-wide window of stratified uniforms staged ahead of the exchange -- every warp stages exactly its own output
  // range after it; >= 0: a window around the CTA's slice with that slack (misses are repaired per warp)
  // default for big slices: a window with 1536 slots of slack on either side (measured on the B200 at N = 2^20: 72.6 G particle-
  // timesteps/s without a window, 80.1 / 81.4 / 82.8 / 71.4 with 512 / 1012 / 1536 / 2048)
  if (uw_req < 0 && g.nb_max >= 4096 && sizeof(Real) == 4) uw_req = 1536;
  g.uw = uw_req >= 0 ? (uw_req + 3) & ~3 : 0;
  g.ucap = uw_req >= 0 ? (g.nb_max + 2 * g.uw + 3) & ~3 : 0;
  g.xstride = G * g.nb_max + 32 * PPT;   // a partly filled lane reads whole 16-byte pairs beyond its particles
  if (sizeof(Real) == 8 && g.nw > 14) { g.uw = 0; g.ucap = 0; }
  g.smem = (size_t)(((G + 1) & ~1) + 40 + 32 + 4 * 32) * sizeof(double) + (size_t)2 * g.ucap * sizeof(unsigned int) +
           (size_t)g.nw * ((size_t)32 * fast_xs(PPT) * sizeof(Real) + (size_t)g.ch * sizeof(unsigned int) + (size_t)g.ch * sizeof(Real));
  return g;
}

struct FastParams {
  FilterDev f;
  int G, ngroups;
  int resample_fn;
  uint4* rec;       // [ngroups][2][G][NUS] LL units: the CTAs' records
  void* xnew;       // [ngroups][xstride] LL elements: uint2 (f32) / uint4 (f64)
  int nb_max;       // largest slice (multiple of PPT)
  int xstride;
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

template <typename Real> __device__ __forceinline__ Real fast_warp_max(Real v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) { const Real t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}

template <typename Model, typename Real, int PPT, int NWMAX>
__global__ void __launch_bounds__(NWMAX * 32, 1) k_fast_bpf(FastParams P) {
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
  const int NW = (int)blockDim.x >> 5;
  const int NT = (int)blockDim.x;
  // shared memory carve-up
  double* s_pre = (double*)smem_raw;                 // [G] the group's records: inclusive sums of the CTAs' weight totals within groups of 32 records
  double* s_grp = s_pre + ((G + 1) & ~1);            // [5][8] per group of 32 records: max, total, sum of squares, state sum, pending state sum
  double* s_max = s_grp + 40;                        // [32] per-warp maxima of the log-weights
  double* s_tot = s_max + 32;                        // [4][32] per-warp sums of e, e^2, e*x and of the resampled x
  unsigned int* s_u = (unsigned int*)(s_tot + 4 * 32);   // [2][ucap] (optional) stratified uniforms staged ahead of the exchange, by observation parity
  unsigned char* s_warp = (unsigned char*)(s_u + 2 * P.ucap);
  constexpr size_t WARP_BYTES = (size_t)32 * XS * sizeof(Real) + (size_t)CH * sizeof(unsigned int) + (size_t)CH * sizeof(Real);
  Real* s_xs = (Real*)(s_warp + (size_t)wid * WARP_BYTES);     // [32][XS] this warp's particles
  unsigned int* s_hd = (unsigned int*)(s_xs + 32 * XS);         // [CH] expansion: epoch << 10 | index into s_xs, at the first slot of a source
  Real* s_out = (Real*)(s_hd + CH);                             // [CH] staging of the chosen x (before that: the warp's window of uniforms)

  uint4* const rec = P.rec + (size_t)group * 2 * G * RL::NUS;
  typedef typename std::conditional<F32, uint2, uint4>::type XEl;   // LL element of x_new
  XEl* xnew = (XEl*)P.xnew + (size_t)group * P.xstride;
  const double INF = __longlong_as_double(0x7FF0000000000000LL), NINF = -INF;
#ifdef BSSM_FAST_TIMING_BUILD
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
#define FAST_TICK(ph) do { if (P.timing && lane == 0) { long long t_ = clock64(); tacc[ph] += t_ - tprev; tprev = t_; } } while (0)
#else
#define FAST_TICK(ph) do { } while (0)
#endif
  unsigned int ep1 = 0, ep2 = 0;   // record / x_new epochs: identical sequences in every CTA of the group
  unsigned int hep = 0;            // head-array epoch of this warp (never reset: stale heads always compare low)
  const int R = (G + 31) >> 5;     // records per lane in the merge
  for (int i = lane; i < CH; i += 32) s_hd[i] = 0u;
  __syncwarp();

  for (int c = group; c < f.C; c += P.ngroups) {
    if (!f.alive[c]) continue;
    const int n = filt_n(f, c);
    int nb = (n + G - 1) / G;
    nb = (nb + PPT - 1) / PPT * PPT;
    const int base = b * nb;                                  // first global particle of this CTA
    const int ws = 32 * PPT;                                  // particles per warp
    const int wbase = min(wid * ws, nb);                      // first particle of this warp within the slice
    const int wcnt = max(0, min(min(ws, nb - wbase), n - (base + wbase)));   // particles of this warp
    const int ibase = base + wid * ws + lane * PPT;           // first global particle of this thread
    const int n_own = max(0, min(PPT, wcnt - lane * PPT));    // owned particles of this thread
    Real par[Model::NPAR];
    Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
    const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
    const int T1 = f.T + 1;
    const int ralg = f.ralg;
    double thr = f.threshold;
    if (thr < 0) thr = (ralg == 0) ? INF : (ralg == 1 ? (double)n : (double)n / 2.0);
    const double log_n = log((double)n);
    const int u_base = max(0, (base - P.uw) & ~3);            // first slot of the CTA's (optional) staged window of uniforms
    // first particle of this warp / of the next one within the slice (clamped to the slice)
    const int lstart = wbase, lnext = min((wid + 1) * ws, nb);

    // what the merge of an exchange leaves in (warp-uniform) registers
    double t_start = 0.0, t_end = 0.0, t_sl = 0.0;   // slot positions where this warp's particles begin / end; slots per unit of e
    double w_start = 0.0;                            // block-local sum of e before this warp
    double loglike = 0.0;                            // meaningful in thread 0 of CTA 0
    int n_resampled = 0, pending_obs = -1;

    double g_M = 0.0;                                // global max of the log-weights of the last exchange
    // ---- one exchange.  phase 0: t = 0 state estimate, 1: observation `obs`, 2: final flush.  In: this warp's max of the log-weights
    //      (phase 1), the warp-inclusive scan of the thread sums of e = exp(lw - ref), the thread's other sums.  Returns dead |
    //      resample << 1, or 4: the weights underflow against `ref`, redo against g_M; leaves t_start / t_end / t_sl / w_start ----
    auto exchange = [&](int phase, int obs, Real mw, double inc_e, Real fq, Real fx, Real fp, Real ref, auto&& overlap) -> int {
      // block totals: per-warp sums -> shared memory -> every warp scans them redundantly (bit-identical in all warps)
      const Real tq = fast_warp_sum<Real>(fq), tx = fast_warp_sum<Real>(fx), tp = fast_warp_sum<Real>(fp);
      if (lane == 31) s_tot[wid] = inc_e;
      if (lane == 0) { s_tot[32 + wid] = (double)tq; s_tot[64 + wid] = (double)tx; s_tot[96 + wid] = (double)tp; s_max[wid] = (double)mw; }
      FAST_TICK(1);   // P1 (propagate, weights, warp reductions)
      __syncthreads();
      const double wt = lane < NW ? *(volatile double*)&s_tot[lane] : 0.0;
      FAST_TICK(8);   // barrier (block totals)
      const double winc = warp_incl_scan_d(wt, lane);
      const double w_end = __shfl_sync(0xffffffffu, winc, wid);                       // block-local sum of e up to and including this warp
      w_start = wid == 0 ? 0.0 : __shfl_sync(0xffffffffu, winc, (wid + 31) & 31);     // ... before this warp: the previous warp's w_end, bit for bit
      ep1++;
      const int par = (int)(ep1 & 1u);
      const int npg = (G + 31) >> 5;                  // groups of 32 records
      if (wid == 0) {
        // the CTA's record, by warp 0
        const double s_b = __shfl_sync(0xffffffffu, winc, 31);
        const Real q_r = fast_warp_sum<Real>(lane < NW ? (Real)s_tot[32 + lane] : (Real)0), x_r = fast_warp_sum<Real>(lane < NW ? (Real)s_tot[64 + lane] : (Real)0),
                   p_r = fast_warp_sum<Real>(lane < NW ? (Real)s_tot[96 + lane] : (Real)0);
        const Real m_r = fast_warp_max<Real>(lane < NW ? (Real)s_max[lane] : Math<Real>::ninf());
        if (G > 1) {
          uint4* dst = rec + (size_t)(par * G + b) * RL::NUS;
          if (F32) {
            if (lane == 0) ll_put_double(dst, s_b, ep1);
            else if (lane == 1) ll_store_v4(dst + 1, __float_as_uint((float)m_r), ep1, __float_as_uint((float)q_r), ep1);
            else if (lane == 2) ll_store_v4(dst + 2, __float_as_uint((float)x_r), ep1, __float_as_uint((float)p_r), ep1);
          } else if (lane < 5) ll_put_double(dst + lane, lane == 0 ? (double)m_r : (lane == 1 ? s_b : (lane == 2 ? (double)q_r : (lane == 3 ? (double)x_r : (double)p_r))), ep1);
        } else if (lane == 0) {
          // a group of one: the CTA's record is the group's
          s_pre[0] = s_b;
          s_grp[0] = (double)m_r; s_grp[8] = s_b; s_grp[16] = (double)q_r; s_grp[24] = (double)x_r; s_grp[32] = (double)p_r;
        }
      }
      FAST_TICK(9);   // block scan, publish
      overlap();      // work that does not depend on the exchange, while the record travels
      FAST_TICK(2);
      // ---- poll the G records, 32 per warp and round: lane l holds record 32 g + l (all its units in flight at once); the warp
      //      leaves the inclusive sums of the weight totals within the group and the group's totals ----
      if (G > 1) {
        const uint4* src = rec + (size_t)par * G * RL::NUS;
        for (int g = wid; g < npg; g += NW) {
          const int j = 32 * g + lane;
          const bool valid = j < G;
          const uint4* rj = src + (size_t)(valid ? j : G - 1) * RL::NUS;
          uint4 v[RL::NU];
          bool ok;
          do {
            ok = true;
#pragma unroll
            for (int i = 0; i < RL::NU; i++) { v[i] = ll_load_v4(rj + i); ok = ok && v[i].y == ep1 && v[i].w == ep1; }
          } while (!ok);
          double sj; Real mj, qj, xj, pj;
          if (F32) {
            sj = ll_get_double(v[0]);
            mj = (Real)__uint_as_float(v[1].x); qj = (Real)__uint_as_float(v[1].z);
            xj = (Real)__uint_as_float(v[2 % RL::NU].x); pj = (Real)__uint_as_float(v[2 % RL::NU].z);
          } else {
            mj = (Real)ll_get_double(v[0]); sj = ll_get_double(v[1]); qj = (Real)ll_get_double(v[2]);
            xj = (Real)ll_get_double(v[3 % RL::NU]); pj = (Real)ll_get_double(v[4 % RL::NU]);
          }
          if (!valid) { sj = 0.0; mj = Math<Real>::ninf(); qj = (Real)0; xj = (Real)0; pj = (Real)0; }
          const double inc = warp_incl_scan_d(sj, lane);
          const Real gm = fast_warp_max<Real>(mj);
          // sums in the summation order of the records (a fixed tree): identical in every CTA
          const double gq = warp_sum_d((double)qj), gx = warp_sum_d((double)xj), gp = warp_sum_d((double)pj);
          if (valid) s_pre[j] = inc;
          if (lane == 31) s_grp[8 + g] = inc;
          if (lane == 0) { s_grp[g] = (double)gm; s_grp[16 + g] = gq; s_grp[24 + g] = gx; s_grp[32 + g] = gp; }
        }
      }
      FAST_TICK(3);   // poll
      __syncthreads();
      const double m0_ = *(volatile double*)&s_grp[0];
      FAST_TICK(10);  // barrier (records)
      // ---- global sums / this CTA's cdf interval: every thread adds up the (at most 8) group totals in the same order, so the
      //      results are bit-identical in every warp of every CTA: no roles, no broadcast, no further barrier ----
      double M = m0_, S = 0.0, Q = 0.0, A_lo = 0.0, A_hi = 0.0;
      {
        const int gl = (b - 1) >> 5, gh = b >> 5;   // groups of records b - 1 and b
        for (int g = 0; g < npg; g++) {
          if (b > 0 && g == gl) A_lo = S + s_pre[b - 1];
          if (g == gh) A_hi = S + s_pre[b];
          const double mg = s_grp[g];
          M = mg > M ? mg : M;
          S += s_grp[8 + g]; Q += s_grp[16 + g];
        }
      }
      g_M = M;
      int dead = 0, resample = 0;
      bool bad = false, empty = false;
      if (phase == 1) {
        bad = (S != S) || (M != M);
        empty = M < -1e8;
        // the weights are exp(lw - ref): if even the best one is tiny the sums have lost their precision (or are 0): redo against M
        if (!bad && !empty && M - (double)ref < (F32 ? -60.0 : -600.0)) return 4;
        dead = (bad || empty) ? 1 : 0;
        // ess < thr  <=>  S^2 < thr * Q  (no division on the critical path)
        resample = dead ? 0 : ((ralg == 0) ? 0 : (ralg == 1 ? 1 : (S * S < thr * Q)));
        if (resample) {
          // slot positions where this warp's particles begin and end.  The end is the next warp's (next CTA's) beginning,
          // formed by the same expression from the same values; a warp that starts at or beyond particle n sits at the end of the slots
          const double nS = (double)n / S;
          t_sl = nS;
          t_start = lstart >= nb ? A_hi * nS : (A_lo + w_start) * nS;
          t_end = lnext >= nb ? A_hi * nS : (A_lo + w_end) * nS;
          if ((long long)base + lstart >= (long long)n) t_start = 2.0 * (double)n;
          if ((long long)base + lnext >= (long long)n) t_end = 2.0 * (double)n;
        }
      }
      FAST_TICK(4);   // merge
      if (b == 0 && tid == 0) {
        // running log-likelihood and the outputs of this observation: one thread, off everybody's critical path
        double SX = 0.0, PEND = 0.0;
        for (int g = 0; g < npg; g++) { SX += s_grp[24 + g]; PEND += s_grp[32 + g]; }
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
              loglike += ((double)ref + log(S) - log_n);
              if (f.loglike_history) f.loglike_history[(size_t)c * f.T + obs] = loglike;
              f.ess[(size_t)c * T1 + obs + 1] = resample ? (double)n : (S * S) / Q;
              if (!resample) f.state_est[(size_t)c * T1 + obs + 1] = SX / S;
            }
          } else {
            f.loglike[c] = loglike; f.n_resampled[c] = n_resampled;
          }
        }
      }
      pending_obs = resample ? obs : -1;
      n_resampled += resample;
      return dead | (resample << 1);
    };

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
    exchange(0, -1, (Real)0, 0.0, (Real)0, px, (Real)0, (Real)0, [] {});
    px = 0;

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
      // ---- P1: propagate + log-weight ----
      for (int tnow = prev_t + 1; tnow <= ot; tnow++) {
        if (zpre_t != tnow - 1) gen_normals(tnow - 1);
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          Real zi[1] = {zpre[k]};
          Model::template transition<Real>(&x[k], par, tnow, zi, nullptr);
        }
      }
      Real e[PPT];   // first the log-weights, then exp(lw - ref)
      Real mloc = Math<Real>::ninf();
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        e[k] = Model::template loglik<Real>(yv, &x[k], par, ot);
        if (k >= n_own) e[k] = Math<Real>::ninf();
        mloc = e[k] > mloc ? e[k] : mloc;
      }
      // a NaN log-weight is not an ordered maximum: it reaches the sums through exp below
      const Real mw = fast_warp_max<Real>(mloc);
      FAST_TICK(11);  // P1 up to the weights
      double exu;            // warp-local exclusive prefix of this thread (unnormalised, relative to the reference)
      double inc_e;
      Real fq, fx;
      // e = exp(lw - ref); thread sums; warp-local inclusive scan of the thread sums
      auto weigh = [&](Real ref) {
        Real fs = 0;
        fq = 0; fx = 0;
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          Real ek = Math<Real>::exp_(e[k] - ref);
          e[k] = ek;
          fs += ek; fq += ek * ek; fx += ek * x[k];
        }
        const double run = (double)fs;
        inc_e = warp_incl_scan_d(run, lane);
        exu = inc_e - run;
      };
      Real ref = Model::template loglik_bound<Real>(par);
      weigh(ref);
      unsigned int w_sys = 0u;
      unsigned int* const su = s_u + (obs & 1) * P.ucap;
      int fl = exchange(1, obs, mw, inc_e, fq, fx, px, ref, [&] {
        if (obs + 1 < f.T) gen_normals(ot);
        if (ralg != 0 && (P.ucap > 0 || P.resample_fn == 1)) {
          if (P.resample_fn == 1) {
            uint4x q0 = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, 0u);
            w_sys = q0.w[0];
          } else {
            // stage the Philox words of the slots this CTA will probably serve: one call per 4 slots
            const int q_end = min((n + 3) >> 2, (u_base + P.ucap) >> 2);
            for (int qd = (u_base >> 2) + tid; qd < q_end; qd += NT) {
              uint4x uq = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, (unsigned int)qd);
              *(uint4*)&su[4 * qd - u_base] = make_uint4(uq.w[0], uq.w[1], uq.w[2], uq.w[3]);
            }
          }
        }
      });
      if (fl == 4) {
        // an outlying observation: every weight is tiny against the bound.  Once more, against the true maximum
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          e[k] = Model::template loglik<Real>(yv, &x[k], par, ot);
          if (k >= n_own) e[k] = Math<Real>::ninf();
        }
        ref = (Real)g_M;
        weigh(ref);
        fl = exchange(1, obs, mw, inc_e, fq, fx, px, ref, [] {});
      }
      px = 0;
      if (fl & 1) break;
      if (!(fl & 2)) continue;

      // ---- P3: closed-form offspring ranges ----
      SlotCounter sc;
      sc.key = key; sc.obs = (unsigned int)obs; sc.fn = P.resample_fn; sc.n = n; sc.w_sys = w_sys;
      sc.s_u = su; sc.u_base = u_base; sc.u_cap = P.ucap;
      // output range of this warp: F at the warp's first particle and at the next warp's (both neighbours evaluate the same values)
      int o_start, o_end;
      {
        const int o = sc.count_slots((lane & 1) ? t_end : t_start);
        o_start = __shfl_sync(0xffffffffu, o, 0);
        o_end = __shfl_sync(0xffffffffu, o, 1);
        if (o_end < o_start) o_end = o_start;   // cannot happen with monotone sums; keeps the ranges sane if it ever did
      }
      if (P.resample_fn != 1 && (o_start - 1 < u_base || o_end + 1 > u_base + P.ucap) && o_end > o_start) {
        // stage the Philox words of exactly this warp's output range (lookups at floor(t) reach one slot below): one call per
        // 4 slots; what does not fit the window (more than CH offspring in the warp) is recomputed per lookup
        const int wb = max(0, (o_start - 1) & ~3);
        unsigned int* const wu = (unsigned int*)s_out;
        const int q_end = min(min((n + 3) >> 2, (o_end + 4) >> 2), (wb + CH) >> 2);
        for (int qd = (wb >> 2) + lane; qd < q_end; qd += 32) {
          uint4x uq = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, (unsigned int)qd);
          *(uint4*)&wu[4 * qd - wb] = make_uint4(uq.w[0], uq.w[1], uq.w[2], uq.w[3]);
        }
        __syncwarp();
        sc.s_u = wu; sc.u_base = wb; sc.u_cap = min(CH, 4 * q_end - wb);
      }
      FAST_TICK(5);   // uniforms of the warp's output range
      // F of this thread's sources (monotone by a running max; clamped into [o_start, o_end])
      int F[PPT];
      {
        const double sl = t_sl;
        const double T0 = t_start + exu * sl;      // slot position just before this thread's first particle
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
            if (lane * PPT + k == wcnt - 1) v = o_end;   // the warp's last particle takes what is left
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
              if (lane * PPT + k == wcnt - 1) v = o_end;
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
      FAST_TICK(6);   // offspring ranges
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
      FAST_TICK(7);   // expansion + LL copy-out
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
    exchange(2, -1, (Real)0, 0.0, (Real)0, (Real)0, px, (Real)0, [] {});
    __syncthreads();
  }    // filters
#ifdef BSSM_FAST_TIMING_BUILD
  if (P.timing && lane == 0 && wid < 16) for (int i = 0; i < 12; i++) P.timing[((size_t)blockIdx.x * 16 + wid) * 12 + i] = tacc[i];
#endif
#undef FAST_TICK
}

}  // namespace bssm
