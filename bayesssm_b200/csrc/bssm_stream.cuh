// bssm_stream.cuh -- streaming bootstrap-filter engine: particles resident in HBM, TWO kernels per
// observation (or, for batches, one cooperative kernel for all of them).  Serves what does not fit the persistent kernel's registers (bssm_fast.cuh): single
// filters beyond ~10^6 particles, large batches [chains x particles], and the particle-sharded
// multi-GPU filter (bssm_shard.cu), where one small record exchange sits between the two kernels.
// Replaces the per-observation loop of .particle_filter_core (R/particle_filter_core.R:123-246)
// for algorithm "BPF" and the resamplers of src/resampling.cpp:5-66 (multinomial / stratified / systematic).
//
//   k_st_step      read x, propagate (one Philox call per 4 particles), log-weight, write x.  A block
//                  walks a contiguous range of tiles; every thread keeps an online (max, sum e,
//                  sum e^2, sum e*x) record of its own particles -- no shuffle, barrier or store in
//                  the tile loop -- and the records meet once, at the block's end.  The LAST block
//                  of a filter to finish (ticket counter) merges the BLOCK records in fixed order:
//                  global max / sum / ESS / log-likelihood / resampling decision, and the exclusive
//                  prefix of the block sums -- the scan offsets of the next kernel.  No separate
//                  finalise launch, no atomics on floating-point data, deterministic.  [8 B / particle]
//   k_st_resample  (steps where resampling fires) read x, recompute the weight, tile-local scan;
//                  INPUT-centric closed-form offspring ranges: a source with cdf value c owns the
//                  output slots [F(c_prev), F(c)), F(c) = #{ i : (i + U_i)/n <= c } -- no search,
//                  no cdf array in memory; chosen states are staged in shared memory and leave
//                  the SM as coalesced vector stores.                            [8 B / particle]
//   k_st_chain     batches: every observation of a group of filters in ONE cooperative launch; a block keeps its
//                  tile range of a filter and alternates the two bodies above (st_step_body, st_resample_body), the
//                  blocks of a filter meet through two words per filter and observation.
//   k_st_mn_*      multinomial resampling (src/resampling.cpp:5-13): the n uniforms drawn already sorted (partial sums
//                  of exponential spacings), served by k_st_resample like the stratified positions.
// Log-weights and the cdf never touch HBM (the weight is recomputed from x and y: a few flops
// against 8 bytes), so the traffic is 16 B per resampled particle-timestep against the 40 B of the
// algorithmic model.
//
// Block boundaries in the output are derived by neighbouring blocks from the same prefix values with
// the same expressions, tile boundaries inside a block are that block's own running sums (clamped into
// the block's interval), so every output slot is written exactly once.  Same Philox keying and tie
// rule (first j with cdf[j] >= pos, clamp) as the other engines.
//
// Storage: row c of x0 / x1 holds the filter's local particles at storage index
// (global index - (goff & ~3)), so Philox quads stay aligned with 16-byte vectors for any shard
// offset goff (goff = 0 when the filter is not sharded).
#pragma once
#include "bssm_common.cuh"
#include "bssm_filter.cuh"
#include "bssm_models.cuh"
#include "bssm_slots.cuh"

namespace bssm {

// threads per block are a template parameter of the kernels (THREADS): 256 for single big filters (fewer block
// records to merge), 128 for batches (smaller barrier domains: +6-8 % there, -3 % on single filters)
constexpr int ST_ERR_CAPACITY = 10; // status: a shard outgrew its storage (BSSM_ERR_CAPACITY)


// layout / cdf descriptor of one filter on this rank, written once per observation
struct StSeg {
  long long goff, ngoff;   // global index of the first local particle: now / after this step's resampling
  int nloc, nnloc;         // local particle count: now / after this step's resampling
  int last, pad;           // 1: this rank holds the tail of the filter
  double abase, aend;      // cdf numerator (relative to the global max) before / after this rank's particles
  double gscale;           // exp(local max - global max)
  double mloc;             // local max (reference of the block prefixes of this rank)
};

struct StRec { double m, s, q, sx, pend, nan, pad0, pad1; };   // per-rank record (sharded runs)
// sharded runs, peer-memory exchange: one slot per (parity of the exchange, sending rank) in every rank's inbox.  The sender
// stores its record into the slot of EVERY rank's inbox over NVLink and releases the slot's sequence word; the receiver polls
// only its own memory.
struct StPeerSlot { StRec rec; unsigned long long seq; unsigned long long pad[7]; };
#define BSSM_PEER_MAX_WORLD 16

// resident threads per SM the two hot kernels are compiled for (register budget = 65536 / this)
#ifndef BSSM_ST_OCC_STEP
#define BSSM_ST_OCC_STEP 1024
#endif
#ifndef BSSM_ST_OCC_RES
#define BSSM_ST_OCC_RES 1024
#endif
struct StreamParams {
  FilterDev f;
  int resample_fn;
  int nt;                  // tiles per filter row (capacity)
  size_t xstride;          // elements per filter row of x0 / x1 (= nt * tile size)
  void *x0, *x1;           // x0: resampled (or initial) particles; x1: propagated particles
  double *blk_m, *blk_s, *blk_q, *blk_x;   // [C][bpc] block records (max; sum e, sum e^2, sum e*x relative to it)
  double* pref;            // [C][bpc + 1] exclusive prefix of the block sums, relative to the local max
  double* bsum;            // [C][bpc] sum of the states written by a block (state estimate after resampling)
  unsigned int* counter;   // [C] tickets
  int* res;                // [2][C] resampling decision of an observation, by parity
  StSeg* seg;              // [2][C] by parity
  int n_glob;              // sharded: global particle count; 0: filt_n
  int sharded, rank, world;
  StRec* rec_local;        // [C] this rank's record
  StRec* rec_all;          // [world][C] gathered records
  StPeerSlot* peer[BSSM_PEER_MAX_WORLD];   // sharded, peer-memory exchange: rank g's inbox [2][world] (peer[rank] is this rank's own); else null
  unsigned long long peer_timeout_ns;      // a peer's record that has not arrived after this long fails the filter (BSSM_ERR_NCCL) instead of hanging the GPU
  unsigned long long peer_seq0;            // sequence number of observation 0's exchange (> 0; 0: the exchange is ncclAllGather + k_st_merge)
  int cap;                 // storage capacity (particles) of a row
  int bpc;                 // blocks per filter (k_st_init, k_st_step and k_st_resample share the block -> tile ranges)
  double log_n;            // log(particle count) when every filter has the same count (else NaN: computed on the device)
  long long* dbg;          // optional [8] clock64 stamps of the merging block (BSSM_ST_TIMING, diagnostics)
  // multinomial resampling (resample_fn == 2): sorted uniforms from exponential spacings
  double* mn_pos;          // [C][xstride] positions of the output slots: U_(0) < U_(1) < ... (order statistics of n uniforms)
  double* mn_tsum;         // [C][mn_nt] sums of the spacings of 1024 slots; after the scan: their exclusive prefix
  double* mn_total;        // [C] sum of all n + 1 spacings
  int mn_nt;               // 1024-slot tiles of spacings per filter row
  int mn_ahead;            // 1: the three arrays above are doubled by the observation's parity and laid out for EVERY observation, ahead of
                           // its resampling decision, on a second stream (one GPU); 0: in line, only where a filter resamples
  // Philox round keys of the launch-wide key word (seed_lo + r * 0x9E3779B9): operands straight from the constant bank
  unsigned int rk0[10];
  // chain-persistent kernel (k_st_chain): per-filter barrier words between the blocks of a filter
  unsigned int* epoch;     // [C] observations whose bookkeeping the merging block has published
  unsigned int* bar2;      // [C] arrivals after resampling steps (monotonic)
};
#ifndef __CUDACC_RTC__
inline void st_fill_round_keys(StreamParams& P) {
  for (int r = 0; r < 10; r++) P.rk0[r] = (unsigned int)P.f.seed + (unsigned int)r * 0x9E3779B9u;
}
#endif
// Philox4x32-10 with the first key word's round keys given (same function as philox4x32_10)
__device__ __forceinline__ uint4x philox4x32_10_rk(const unsigned int (&rk0)[10], uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ rk0[r], n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k1 += 0xBB67AE85u;
  }
  uint4x o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
  return o;
}

// explicitly rounded a + p * g: the same boundary value in every tile / rank that evaluates it
__device__ __forceinline__ double st_bound(double a, double p, double g) { return __dadd_rn(a, __dmul_rn(p, g)); }

template <typename Real, int PPT> struct StVec {
  static __device__ __forceinline__ void load(Real* x, const Real* p) {
    if constexpr (sizeof(Real) == 4) {
#pragma unroll
      for (int h = 0; h < PPT / 4; h++) { float4 v = *(const float4*)(p + 4 * h); x[4 * h] = v.x; x[4 * h + 1] = v.y; x[4 * h + 2] = v.z; x[4 * h + 3] = v.w; }
    } else {
#pragma unroll
      for (int h = 0; h < PPT / 2; h++) { double2 v = *(const double2*)(p + 2 * h); x[2 * h] = v.x; x[2 * h + 1] = v.y; }
    }
  }
  static __device__ __forceinline__ void store(Real* p, const Real* x) {
    if constexpr (sizeof(Real) == 4) {
#pragma unroll
      for (int h = 0; h < PPT / 4; h++) *(float4*)(p + 4 * h) = make_float4(x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3]);
    } else {
#pragma unroll
      for (int h = 0; h < PPT / 2; h++) *(double2*)(p + 2 * h) = make_double2(x[2 * h], x[2 * h + 1]);
    }
  }
};

// layout an observation's kernels work on, and the split of its tiles over the blocks of a filter:
// block j owns the contiguous tiles [j * tpb, min(ntc, (j + 1) * tpb)), nb blocks are active
struct StLayout { long long goff; int nloc, lead, ntc, tpb, nb; };
template <int TS>
__device__ __forceinline__ StLayout st_make_layout(long long goff, int nloc, int bpc) {
  StLayout L;
  L.goff = goff; L.nloc = nloc;
  L.lead = (int)(goff & 3);
  L.ntc = max(1, (nloc + L.lead + TS - 1) / TS);
  L.tpb = (L.ntc + bpc - 1) / bpc;
  L.nb = (L.ntc + L.tpb - 1) / L.tpb;
  return L;
}
// the previous observation's descriptor, after its resampling if it fired
template <int TS>
__device__ __forceinline__ StLayout st_layout_in(const StreamParams& P, int c, int obs) {
  const int pp = (obs + 1) & 1;
  const StSeg& sp = P.seg[pp * P.f.C + c];
  const int rprev = P.res[pp * P.f.C + c];
  return st_make_layout<TS>(rprev ? sp.ngoff : sp.goff, rprev ? sp.nnloc : sp.nloc, P.bpc);
}

// ---- global part of the per-observation bookkeeping (R/particle_filter_core.R:189-224) ----
// recs: `world` records in rank order.  Every rank evaluates the same expressions on the same records.
// The work is split into three independent roles so that the merging block of k_st_step can run them on
// three different warps at once (each is a chain of dependent fp64 operations executed by a single thread,
// and fp64 latency is what the tail of the kernel is made of): ST_DECIDE publishes what the resampling
// kernel needs, ST_LOGLIKE the running log-likelihood, ST_ESTIMATES the ESS / state-estimate outputs.
enum { ST_DECIDE = 1, ST_LOGLIKE = 2, ST_ESTIMATES = 4, ST_ALL_ROLES = 7 };
static __device__ __noinline__ void st_global(const StreamParams& P, int c, int obs, const StRec* recs, int rstride, int world, int rank,
                                              long long goff, int nloc, int roles) {
  const FilterDev& f = P.f;
  const double NINF = -__longlong_as_double(0x7FF0000000000000LL);
  const int C = f.C, pc = obs & 1, pp = (obs + 1) & 1;
  const int n = P.n_glob ? P.n_glob : filt_n(f, c);
  const int T1 = f.T + 1;
  const int rprev = P.res[pp * C + c];
  double M = NINF;
  int nan = 0;
  for (int g = 0; g < world; g++) { const StRec& r = recs[(size_t)g * rstride]; M = r.m > M ? r.m : M; nan |= (r.nan != 0.0) || (r.m != r.m); }
  double S = 0.0, Q = 0.0, SX = 0.0, PEND = 0.0, abase = 0.0, aend = 0.0, gscale = 0.0;
  for (int g = 0; g < world; g++) {
    const StRec& r = recs[(size_t)g * rstride];
    double sc = 0.0;
    if (r.m == M) sc = (M == NINF) ? 0.0 : 1.0;
    else if (r.m != NINF) sc = exp(r.m - M);
    if (g == rank) { abase = S; gscale = sc; }
    S = st_bound(S, r.s, sc);
    if (g == rank) aend = S;
    Q += r.q * sc * sc; SX += r.sx * sc; PEND += r.pend;
  }
  const bool bad = nan || (S != S) || (SX != SX) || (M != M);
  const bool empty = M < -1e8;
  const int ralg = f.ralg;
  double thr = f.threshold;
  if (thr < 0) thr = (ralg == 0) ? -NINF : (ralg == 1 ? (double)n : 0.5 * (double)n);
  // ess < thr  <=>  S^2 < thr * Q  (no division on the path the next kernel waits for)
  int resample = (bad || empty) ? 0 : ((ralg == 0) ? 0 : (ralg == 1 ? 1 : (S * S < thr * Q)));

  if (roles & ST_DECIDE) {
    StSeg sg;
    sg.goff = goff; sg.nloc = nloc; sg.ngoff = goff; sg.nnloc = nloc; sg.last = (rank == world - 1); sg.pad = 0;
    sg.abase = abase; sg.aend = aend; sg.gscale = gscale; sg.mloc = recs[(size_t)rank * rstride].m;
    if (bad) { f.status[c] = 3; f.alive[c] = 0; }                  // NaN weight somewhere: R's `if (NA)` error
    else if (empty) { f.early_exit[c] = 1; f.alive[c] = 0; }       // all(lw < -1e8): R/particle_filter_core.R:189-202
    else if (resample) {
      f.n_resampled[c] += 1;
      if (world > 1) {               // rank g's share of the output slots: [F(A_g / S), F(A_{g+1} / S))
        SlotCounter sc;
        sc.key = make_key(f.seed, f.run_id[c], f.stream[c]); sc.obs = (unsigned int)obs; sc.fn = P.resample_fn; sc.n = n;
        sc.s_u = nullptr; sc.u_base = 0; sc.u_cap = 0; sc.w_sys = 0u;
        if (P.resample_fn == 1) { uint4x q0 = noise_quad(sc.key, (unsigned int)obs, TAG_RESAMP_U, 0u, 0u); sc.w_sys = q0.w[0]; }
        // every rank checks every rank's share, so a capacity overflow stops the whole group consistently
        double run = 0.0;
        int o_prev = 0, overflow = 0;
        for (int g = 0; g < world; g++) {
          const StRec& r = recs[(size_t)g * rstride];
          double scg = 0.0;
          if (r.m == M) scg = 1.0;
          else if (r.m != NINF) scg = exp(r.m - M);
          run = st_bound(run, r.s, scg);
          const int o_next = (g == world - 1) ? n : sc.count_le(run / S);
          if (o_next - o_prev > P.cap) overflow = 1;   // a row stores cap + 4 elements and the share starts at most 3 in: it fits
          if (g == rank) { sg.ngoff = o_prev; sg.nnloc = o_next - o_prev; }
          o_prev = o_next;
        }
        if (overflow) { f.status[c] = ST_ERR_CAPACITY; f.alive[c] = 0; resample = 0; }
      }
    }
    f.M[c] = M; f.S[c] = S;
    P.seg[pc * C + c] = sg;
    P.res[pc * C + c] = resample;
  }
  if (roles & ST_LOGLIKE) {
    if (!bad) {
      const double ll = empty ? NINF : f.loglike[c] + (M + log(S) - ((P.log_n == P.log_n) ? P.log_n : log((double)n)));
      f.loglike[c] = ll;
      if (f.loglike_history) f.loglike_history[(size_t)c * f.T + obs] = ll;
    }
  }
  if (roles & ST_ESTIMATES) {
    if (obs == 0) f.ess[(size_t)c * T1] = (double)n;
    if (rprev) f.state_est[(size_t)c * T1 + obs] = PEND / (double)n;   // initial particles (obs = 0) or the previous resampling
    if (!bad && !empty) {
      f.ess[(size_t)c * T1 + obs + 1] = resample ? (double)n : (S * S) / Q;
      if (!resample) f.state_est[(size_t)c * T1 + obs + 1] = SX / S;
    }
  }
}

// ---- local merge of the block records by the last block of a filter (all threads of the block) ----
// Produces this rank's record and the exclusive prefix of the block sums (relative to the local max).
// One SM does this while the others idle, so there is little of it (one record per BLOCK, not per tile:
// at most a few rounds) and it is laid out for that SM's memory pipeline: SoA arrays, lane = record
// (coalesced), every load of a round issued before the first use.  Fixed structure => deterministic.
// Plain loads: the records were published with fence + ticket and are read after ticket + fence, and no
// line of them was in this SM's L1 before.
template <typename Real, int ST_THREADS>
static __device__ __forceinline__ void st_local_merge(const StreamParams& P, int c, int nb, int nb_pending, double* s_red /*[4][ST_NW]*/, StRec& out) {
  constexpr int ST_NW = ST_THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const double NINF = -__longlong_as_double(0x7FF0000000000000LL);
  const size_t row = (size_t)c * P.bpc;
  const double* pm = P.blk_m + row;
  const double* ps = P.blk_s + row;
  const double* pq = P.blk_q + row;
  const double* px = P.blk_x + row;
  const double* bsum = P.bsum + row;
  double* pref = P.pref + (size_t)c * (P.bpc + 1);
  const int seg = ((nb + ST_NW - 1) / ST_NW + 31) & ~31;      // records per warp, multiple of 32
  const int j0 = min(nb, wid * seg), j1 = min(nb, j0 + seg);
  if (P.dbg && tid == 0) P.dbg[4] = clock64();
  // pending state sum of the previous resampling (or of the initial particles): its own record count
  double lp = 0.0;
  for (int j = tid; j < nb_pending; j += ST_THREADS) lp += bsum[j];
  double m = NINF;
#pragma unroll 4
  for (int j = j0 + lane; j < j1; j += 32) { const double mj = pm[j]; m = mj > m ? mj : m; }   // NaN maxima are caught through s
  m = warp_max_d(m);
  if (lane == 0) s_red[wid] = m;
  __syncthreads();
  m = warp_max_d(lane < ST_NW ? s_red[lane] : NINF);
  __syncthreads();
  if (P.dbg && tid == 0) P.dbg[5] = clock64();
  double lq = 0.0, lx = 0.0, carry = 0.0;
#pragma unroll 2
  for (int jb = j0; jb < j1; jb += 32) {
    const int j = jb + lane;
    const bool on = j < j1;
    const double mj = on ? pm[j] : NINF, sj = on ? ps[j] : 0.0;
    const double qj = on ? pq[j] : 0.0, xj = on ? px[j] : 0.0;
    double sc = 0.0;
    if (!(mj == NINF || m == NINF)) sc = (double)Math<Real>::exp_((Real)(mj - m));   // throughput precision: SFU exp
    const double v = sj * sc;
    lq += qj * sc * sc; lx += xj * sc;
    const double inc = warp_incl_scan_d(v, lane);
    if (on) pref[j] = carry + (inc - v);            // exclusive prefix relative to this warp's segment
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (P.dbg && tid == 0) P.dbg[6] = clock64();
  const double tq = warp_sum_d(lq), tx = warp_sum_d(lx), tp = warp_sum_d(lp);
  if (lane == 0) { s_red[wid] = carry; s_red[ST_NW + wid] = tq; s_red[2 * ST_NW + wid] = tx; s_red[3 * ST_NW + wid] = tp; }
  __syncthreads();
  double woff = 0.0, S = 0.0, Q = 0.0, SX = 0.0, PD = 0.0;
#pragma unroll
  for (int w = 0; w < ST_NW; w++) {                 // fixed order
    if (w == wid) woff = S;
    S += s_red[w]; Q += s_red[ST_NW + w]; SX += s_red[2 * ST_NW + w]; PD += s_red[3 * ST_NW + w];
  }
  if (P.dbg && tid == 0) P.dbg[7] = clock64();
  if (wid > 0) {
#pragma unroll 4
    for (int j = j0 + lane; j < j1; j += 32) pref[j] += woff;     // own values: the same thread wrote them
  }
  if (tid == 0) pref[nb] = S;
  out.m = m; out.s = S; out.q = Q; out.sx = SX; out.pend = PD; out.nan = (S != S) ? 1.0 : 0.0; out.pad0 = 0; out.pad1 = 0;
  __syncthreads();
}

// ---- setup: descriptors of "observation -1" (so that observation 0 reads the initial particles from x0) ----
static __global__ void k_st_setup(StreamParams P, long long goff0, int nloc0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.f.C) return;
  const int n = P.n_glob ? P.n_glob : filt_n(P.f, c);
  StSeg sg;
  sg.goff = sg.ngoff = P.sharded ? goff0 : 0;
  sg.nloc = sg.nnloc = P.sharded ? nloc0 : n;
  sg.last = P.rank == P.world - 1; sg.pad = 0; sg.abase = 0; sg.aend = 0; sg.gscale = 1; sg.mloc = 0;
  P.seg[1 * P.f.C + c] = sg; P.seg[c] = sg;
  P.res[1 * P.f.C + c] = 1; P.res[c] = 0;
  P.counter[c] = 0u;
  if (P.epoch) { P.epoch[c] = 0u; P.bar2[c] = 0u; }
#ifdef BSSM_ST_CHAIN_TIMING
  if (c == 0) g_st_chain_wait = 0ull;
#endif
}

// ---- init (R/particle_filter_core.R:76-116): x0 <- init_fn, block sums for the t = 0 state estimate ----
template <typename Model, typename Real, int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS) k_st_init(StreamParams P) {
  constexpr int ST_THREADS = THREADS, ST_NW = THREADS / 32;
  constexpr int TS = ST_THREADS * PPT;
  __shared__ double s_red[ST_NW];
  const FilterDev& f = P.f;
  const int c = blockIdx.x / P.bpc, j = blockIdx.x % P.bpc, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (!f.alive[c]) return;
  const StLayout L = st_layout_in<TS>(P, c, 0);
  if (j >= L.nb) return;
  Real par[Model::NPAR];
  Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
  const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  double sum0 = 0.0;
  for (int tile = j * L.tpb; tile < min(L.ntc, (j + 1) * L.tpb); tile++) {
    const int sbase = tile * TS + tid * PPT;
    const long long g0 = L.goff - L.lead + sbase;
    Real x[PPT];
#pragma unroll
    for (int h = 0; h < PPT / 4; h++) {
      uint4x qd = noise_quad(key, T_INIT, TAG_INIT_Z, 0u, (unsigned int)((g0 + 4 * h) >> 2));
      Real zz[4];
      Math<Real>::box_muller4(qd.w, zz);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        Real xi[1]; Real zi[1] = {zz[k]};
        Model::template init<Real>(xi, par, zi, nullptr);
        const long long g = g0 + 4 * h + k;
        const bool valid = g >= L.goff && g < L.goff + L.nloc;
        x[4 * h + k] = valid ? xi[0] : (Real)0;
        sum0 += (double)x[4 * h + k];
      }
    }
    StVec<Real, PPT>::store((Real*)P.x0 + (size_t)c * P.xstride + sbase, x);
  }
  double v = warp_sum_d(sum0);
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = warp_sum_d(lane < ST_NW ? s_red[lane] : 0.0);
    if (lane == 0) P.bsum[(size_t)c * P.bpc + j] = t;
  }
}

// per-thread asynchronous prefetch of the next tile's 32 bytes (cp.async, no registers held while in flight).
// Layout [2 halves][ST_THREADS] x 16 B: conflict-free 128-bit shared loads.
#ifndef BSSM_EMU
__device__ __forceinline__ void st_cp_async16(void* smem, const void* gmem) {
  const unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void st_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void st_cp_async_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else   // CPU logic test (tests/simt_emu.h): the copy completes at once
inline void st_cp_async16(void* smem, const void* gmem) { memcpy(smem, gmem, 16); }
inline void st_cp_async_commit() {}
inline void st_cp_async_wait() {}
#endif
template <typename Real, int PPT, int ST_THREADS>
__device__ __forceinline__ void st_prefetch(uint4* buf /*[2][ST_THREADS]*/, const Real* g) {
  static_assert(PPT * sizeof(Real) == 32, "32 bytes per thread and tile");
  st_cp_async16(&buf[threadIdx.x], g);
  st_cp_async16(&buf[ST_THREADS + threadIdx.x], (const char*)g + 16);
  st_cp_async_commit();
}
template <typename Real, int PPT, int ST_THREADS>
__device__ __forceinline__ void st_take(Real* x, const uint4* buf) {
  st_cp_async_wait();
  const uint4 a = buf[threadIdx.x], b = buf[ST_THREADS + threadIdx.x];
  if constexpr (sizeof(Real) == 4) {
    x[0] = __uint_as_float(a.x); x[1] = __uint_as_float(a.y); x[2] = __uint_as_float(a.z); x[3] = __uint_as_float(a.w);
    x[4] = __uint_as_float(b.x); x[5] = __uint_as_float(b.y); x[6] = __uint_as_float(b.z); x[7] = __uint_as_float(b.w);
  } else {
    x[0] = __longlong_as_double(((long long)a.y << 32) | a.x); x[1] = __longlong_as_double(((long long)a.w << 32) | a.z);
    x[2] = __longlong_as_double(((long long)b.y << 32) | b.x); x[3] = __longlong_as_double(((long long)b.w << 32) | b.z);
  }
}
template <typename Real> __device__ __forceinline__ Real st_warp_sum(Real v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename Real> __device__ __forceinline__ Real st_warp_max(Real v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) { Real t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}

// Programmatic dependent launch (sm_90+): k_st_step and k_st_resample of a one-GPU filter are chained with
// cudaLaunchAttributeProgrammaticStreamSerialization.  A kernel lets its successor's blocks become resident as soon as
// its own blocks make room (launch_dependents, first thing), and reads nothing the predecessor wrote before
// griddepcontrol.wait -- which returns when the predecessor grid has completed and its writes are visible.  What sits above
// the wait (parameter set-up, Philox key, the observation) depends only on data that was final two launches earlier.
#if !defined(BSSM_EMU)
__device__ __forceinline__ void st_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void st_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#else
inline void st_pdl_launch_dependents() {}
inline void st_pdl_wait() {}
#endif

#ifdef BSSM_ST_CHAIN_TIMING
__device__ unsigned long long g_st_chain_wait;   // diagnostics build: cycles spent waiting for the merging block
#endif
// ---- barriers between the blocks of ONE filter (chain-persistent kernel k_st_chain; the blocks are co-resident) ----
// One fence per side and no more: the publisher's release (store or reduction) is the only MEMBAR, the waiter's acquire load brings
// the L1 invalidation with it (LDG.STRONG.GPU + CCTL.IVALL) -- it is thread 0's, but the L1 is the SM's, and the other threads of
// the block read nothing between that load and the block barrier that follows it.  (A __threadfence() by every thread on both
// sides, the first version, was a MEMBAR.SC per warp and barrier: profiles/r2_ab_chain_on_off.txt.)
#ifndef BSSM_EMU
__device__ __forceinline__ unsigned int st_ld_acquire(const unsigned int* p) {
  unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_st_release(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_red_release(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
#else
inline unsigned int st_ld_acquire(const unsigned int* p) { emu_poll_yield(); return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_st_release(unsigned int* p, unsigned int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void st_red_release(unsigned int* p, unsigned int v) { __atomic_fetch_add(p, v, __ATOMIC_RELEASE); }
#endif
// publish: everything the block wrote (any thread) before the call is visible to a block that has seen the value
__device__ __forceinline__ void st_chain_publish(unsigned int* p, unsigned int v) {
  __syncthreads();
  if (threadIdx.x == 0) st_st_release(p, v);
}
// wait until the word reaches v; afterwards every thread of the block reads what the publisher wrote (the descriptors live at the
// same addresses every other observation: the acquire load drops this SM's stale L1 lines)
__device__ __forceinline__ void st_chain_wait(const unsigned int* p, unsigned int v) {
  if (threadIdx.x == 0) { while ((int)(st_ld_acquire(p) - v) < 0) {} }
  __syncthreads();
}
// arrive + wait on a monotonic counter (target = arrivals of all rounds so far)
__device__ __forceinline__ void st_chain_arrive_wait(unsigned int* p, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) { st_red_release(p, 1u); while ((int)(st_ld_acquire(p) - target) < 0) {} }
  __syncthreads();
}

// ---- particle-sharded filter: the all-gather of the ranks' records, fused into the kernel that produces them ----------------
// One thread (the merging block's thread 0) stores the rank's 48 bytes into its slot of every rank's inbox -- peer memory mapped
// with CUDA IPC, the stores travel over NVLink / NVSwitch -- fences system-wide once, stores the slots' sequence words, and then polls
// the `world` sequence words of its OWN inbox (local HBM / L2) before copying the records to rec_all.  Slots alternate with the
// parity of the sequence number: a rank can only reach exchange k + 2 after every rank has published exchange k + 1, i.e. after
// every rank's kernel of exchange k has finished reading.  No kernel of another GPU has to be resident for this one to finish (a
// peer's record arrives when that peer's own step kernel reaches its tail), so there is no co-scheduling requirement.  A peer that
// never arrives is a failed job: after P.peer_timeout_ns (30 s; $BSSM_PEER_TIMEOUT_MS) the filter is marked BSSM_ERR_NCCL and dead instead of hanging the GPU.
#ifndef BSSM_EMU
__device__ __forceinline__ unsigned long long st_ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long st_globaltimer() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
static __device__ __noinline__ bool st_peer_allgather(const StreamParams& P, int c, const StRec& r, unsigned long long seq) {
  const int par = (int)(seq & 1ull), W = P.world;
  for (int g = 0; g < W; g++) {
    volatile double* d = (volatile double*)&P.peer[g][(size_t)par * W + P.rank].rec;
    d[0] = r.m; d[1] = r.s; d[2] = r.q; d[3] = r.sx; d[4] = r.pend; d[5] = r.nan;
  }
  __threadfence_system();   // ONE system-scope fence between the records and the sequence words (a release store per peer would be `world` fences in a row)
  for (int g = 0; g < W; g++) st_st_relaxed_sys(&P.peer[g][(size_t)par * W + P.rank].seq, seq);
  const unsigned long long t0 = st_globaltimer();
  for (int g = 0; g < W; g++) {
    const StPeerSlot* src = &P.peer[P.rank][(size_t)par * W + g];
    while (st_ld_acquire_sys(&src->seq) < seq) {
      if (st_globaltimer() - t0 > P.peer_timeout_ns) return false;
    }
    const volatile double* q = (const volatile double*)&src->rec;
    StRec o; o.m = q[0]; o.s = q[1]; o.q = q[2]; o.sx = q[3]; o.pend = q[4]; o.nan = q[5]; o.pad0 = o.pad1 = 0.0;
    P.rec_all[(size_t)g * P.f.C + c] = o;
  }
  return true;
}
#else
inline bool st_peer_allgather(const StreamParams&, int, const StRec&, unsigned long long) { return false; }   // emulated ranks run one after the other
#endif
static __device__ __forceinline__ void st_peer_failed(const StreamParams& P, int c) {
  P.f.status[c] = 11 /*BSSM_ERR_NCCL: the exchange between the ranks failed*/; P.f.alive[c] = 0;
}

// shared memory of the two per-observation bodies (the chain-persistent kernel overlays them)
template <typename Real, int THREADS> struct StStepSmem {
  uint4 pf[2][2 * THREADS];
  Real w[4][THREADS / 32];        // per-warp records at the end of the block
  double red[4 * (THREADS / 32)];
};
template <typename Real, int PPT, int THREADS> struct StResSmem {
  static constexpr int CAP = THREADS * ((PPT * 5 / 4 + 1) & ~1);   // staging capacity: a quarter beyond the tile (more offspring take further chunks)
  // one buffer, two lives: the staged stratified uniforms (raw Philox words) of the tile while the offspring
  // ranges are computed, then the staged outputs of a chunk (multinomial: the staged positions, doubles)
  __align__(16) unsigned char uo[CAP * 8];
  __align__(16) unsigned int head[CAP + THREADS];   // expansion: (source index << 16 | address of its x) at the first slot of a source (+ a spare word per thread)
  uint4 pf[2][2 * THREADS];
  double red[THREADS / 32], bs[THREADS / 32];
  int wf[THREADS / 32];
  unsigned int wh[THREADS / 32];
  int mn[2];
};
enum { ST_GO = 0, ST_LEAVE = 1 };   // body results: carry on / this block has nothing more to do for the filter (dead filter, no tiles)

// ---- K_A: propagate + log-weight + tile / block partials; the last block of a filter merges ----
// Block (c, j) walks the contiguous tiles [j * tpb, (j + 1) * tpb) of filter c; the next tile's particles
// are in flight (cp.async) while the current tile is computed, and the block pays the descriptor loads,
// the parameter set-up and the fence + ticket once, not once per tile.
// PERSIST: called once per observation by the chain-persistent kernel -- the blocks of a filter meet at the end of the body
// (the merging block publishes the observation's bookkeeping, the others wait for it) instead of at a kernel boundary.
template <typename Model, typename Real, int PPT, int THREADS, bool PERSIST>
__device__ __forceinline__ int st_step_body(const StreamParams& P, int obs, int c, int j, const Real* par, const NoiseKey& key,
                                            StStepSmem<Real, THREADS>& sm) {
  constexpr int ST_THREADS = THREADS, ST_NW = THREADS / 32;
  // serves 1-D models with one normal per init / transition and no uniforms; checked on the host (stream_supported),
  // because NVRTC instantiates these kernels for every user model whatever its shape
  static_assert(PPT % 4 == 0, "one Philox call serves 4 particles");
  constexpr int TS = ST_THREADS * PPT;
  uint4 (&s_pf)[2][2 * ST_THREADS] = sm.pf;
  Real (&s_w)[4][ST_NW] = sm.w;
  double* const s_red = sm.red;
  const FilterDev& f = P.f;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // independent of the previous launch: the observation
  const int ot = f.obs_times ? f.obs_times[obs] : obs + 1;
  const int prev_t = obs == 0 ? 0 : (f.obs_times ? f.obs_times[obs - 1] : obs);
  double yv[4] = {0, 0, 0, 0};
  for (int k = 0; k < f.dy && k < 4; k++) yv[k] = f.y[(size_t)obs * f.dy + k];
  if constexpr (!PERSIST) st_pdl_wait();
  const long long t_start = (!PERSIST && P.dbg) ? clock64() : 0;
  if (!f.alive[c]) return ST_LEAVE;
  const StLayout L = st_layout_in<TS>(P, c, obs);
  if (j >= L.nb) return ST_LEAVE;
  const int pp = (obs + 1) & 1;
  const int rprev = P.res[pp * f.C + c];
  const int t0 = j * L.tpb, t1 = min(L.ntc, t0 + L.tpb);
  const Real* xin = (const Real*)(rprev ? P.x0 : P.x1) + (size_t)c * P.xstride + tid * PPT;
  Real* xout = (Real*)P.x1 + (size_t)c * P.xstride + tid * PPT;
  st_prefetch<Real, PPT, ST_THREADS>(s_pf[0], xin + (size_t)t0 * TS);
  // this thread's record: running max and sums (relative to it) over its particles of all the block's tiles
  // (a NaN log-weight never raises the max and turns exp(NaN - ref) into NaN: it poisons the sum by itself --
  // R's `if (NA)` error, reported as BSSM_ERR_NAN_WEIGHT)
  Real mT = Math<Real>::ninf(), sT = 0, qT = 0, xT = 0;
  F2 s2 = f2_make(0.f, 0.f), q2 = s2, x2 = s2;     // throughput precision: the three sums as packed pairs (even / odd particles)

  for (int tile = t0; tile < t1; tile++) {
    const int pb = (tile - t0) & 1;
    Real x[PPT];
    st_take<Real, PPT, ST_THREADS>(x, s_pf[pb]);
    if (tile + 1 < t1) st_prefetch<Real, PPT, ST_THREADS>(s_pf[pb ^ 1], xin + (size_t)(tile + 1) * TS);
    const int sbase = tile * TS + tid * PPT;
    const long long g0 = L.goff - L.lead + sbase;
    const int k_lo = (int)max(0LL, min((long long)PPT, L.goff - g0));
    const int k_hi = (int)max(0LL, min((long long)PPT, L.goff + L.nloc - g0));
    const bool ragged = k_lo > 0 || k_hi < PPT;   // only the first / last threads of a shard hold padding lanes
    if (ragged) {
#pragma unroll
      for (int k = 0; k < PPT; k++) if (k < k_lo || k >= k_hi) x[k] = (Real)0;   // padding lanes stay finite
    }
    for (int tnow = prev_t + 1; tnow <= ot; tnow++) {
#pragma unroll
      for (int h = 0; h < PPT / 4; h++) {
        uint4x qd = noise_quad(key, (unsigned int)(tnow - 1), TAG_TRANS_Z, 0u, (unsigned int)((g0 + 4 * h) >> 2));
        Real zz[4];
        Math<Real>::box_muller4(qd.w, zz);
        if constexpr (sizeof(Real) == 4 && ModelPacked<Model>::value) {
#pragma unroll
          for (int k = 0; k < 4; k += 2)
            f2_get(Model::transition2(f2_make(x[4 * h + k], x[4 * h + k + 1]), par, tnow, f2_make(zz[k], zz[k + 1])), x[4 * h + k], x[4 * h + k + 1]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            Real zi[1] = {zz[k]};
            Model::template transition<Real>(&x[4 * h + k], par, tnow, zi, nullptr);
          }
        }
      }
    }
    StVec<Real, PPT>::store(xout + (size_t)tile * TS, x);
    Real e[PPT];
    Real mloc = Math<Real>::ninf();
#pragma unroll
    for (int k = 0; k < PPT; k += 2) {
      if constexpr (sizeof(Real) == 4 && ModelPacked<Model>::value) f2_get(Model::loglik2(yv, f2_make(x[k], x[k + 1]), par, ot), e[k], e[k + 1]);
      else { e[k] = Model::template loglik<Real>(yv, &x[k], par, ot); e[k + 1] = Model::template loglik<Real>(yv, &x[k + 1], par, ot); }
    }
    if (ragged) {
#pragma unroll
      for (int k = 0; k < PPT; k++) if (k < k_lo || k >= k_hi) e[k] = Math<Real>::ninf();
    }
#pragma unroll
    for (int k = 0; k < PPT; k++) mloc = Math<Real>::max_(mloc, e[k]);
    // Thread-local online accumulation: e relative to this THREAD's running max (rescaled on the rare tiles that
    // raise it).  No shuffle, no barrier, no store of partials in this loop: the sums meet once, at the block's end.
    if (mloc > mT) {
      const Real r = (mT == Math<Real>::ninf()) ? (Real)0 : Math<Real>::exp_(mT - mloc);
      if constexpr (sizeof(Real) == 4) {
        const F2 r2 = f2_make((float)r, (float)r);
        s2 = f2_mul(s2, r2); q2 = f2_mul(q2, f2_mul(r2, r2)); x2 = f2_mul(x2, r2);
      } else { sT *= r; qT *= r * r; xT *= r; }
      mT = mloc;
    }
    {
      const Real mref = (mT == Math<Real>::ninf()) ? (Real)0 : mT;
      if constexpr (sizeof(Real) == 4) {
        // two particles per instruction: exponent arguments and the three sums as packed fp32 pairs (lane sums meet after the loop)
        const F2 nm2 = f2_make(-mref, -mref), l2e = f2_make(1.4426950408889634f, 1.4426950408889634f);
#pragma unroll
        for (int k = 0; k < PPT; k += 2) {
          float a0, a1;
          f2_get(f2_mul(f2_add(f2_make(e[k], e[k + 1]), nm2), l2e), a0, a1);
          const F2 ek2 = f2_make(Math<float>::ex2_(a0), Math<float>::ex2_(a1));
          s2 = f2_add(s2, ek2); q2 = f2_fma(ek2, ek2, q2); x2 = f2_fma(ek2, f2_make(x[k], x[k + 1]), x2);
        }
      } else {
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          const Real ek = Math<Real>::exp_(e[k] - mref);
          sT += ek; qT += ek * ek; xT += ek * x[k];
        }
      }
    }
  }
  if constexpr (sizeof(Real) == 4) {
    float a, b;
    f2_get(s2, a, b); sT = a + b; f2_get(q2, a, b); qT = a + b; f2_get(x2, a, b); xT = a + b;
  }
  // block record: thread records -> warp records (shuffles) -> block record (fixed order)
  {
    const Real mw = st_warp_max<Real>(mT);
    Real sc = (Real)1;                                       // empty thread record: zeros or the NaN marker pass through
    if (mT != Math<Real>::ninf()) sc = Math<Real>::exp_(mT - mw);
    const Real fs = st_warp_sum<Real>(sT * sc), fq = st_warp_sum<Real>(qT * sc * sc), fx = st_warp_sum<Real>(xT * sc);
    if (lane == 0) { s_w[0][wid] = mw; s_w[1][wid] = fs; s_w[2][wid] = fq; s_w[3][wid] = fx; }
  }
  __syncthreads();
  Real mB = Math<Real>::ninf(), sB = 0, qB = 0, xB = 0;
  if (tid == 0) {
#pragma unroll
    for (int w = 0; w < ST_NW; w++) mB = s_w[0][w] > mB ? s_w[0][w] : mB;
#pragma unroll
    for (int w = 0; w < ST_NW; w++) {
      const Real m = s_w[0][w];
      const Real sc = (m == Math<Real>::ninf()) ? (Real)1 : Math<Real>::exp_(m - mB);
      sB += s_w[1][w] * sc; qB += s_w[2][w] * sc * sc; xB += s_w[3][w] * sc;
    }
  }
  // one ticket per block; the last block of the filter merges.  The barrier-reduction makes the outcome a
  // block-uniform value the compiler can see
  int mine = 0;
  if (tid == 0) {
    const size_t brow = (size_t)c * P.bpc + j;
    P.blk_m[brow] = (double)mB; P.blk_s[brow] = (double)sB; P.blk_q[brow] = (double)qB; P.blk_x[brow] = (double)xB;
    __threadfence();
    const unsigned int ticket = atomicAdd(&P.counter[c], 1u);
    mine = (ticket == (unsigned int)(L.nb - 1));
  }
  if (!__syncthreads_or(mine)) {
    if constexpr (PERSIST) {
#ifdef BSSM_ST_CHAIN_TIMING
      const long long tw_ = clock64();
#endif
      st_chain_wait(&P.epoch[c], (unsigned int)(obs + 1));
#ifdef BSSM_ST_CHAIN_TIMING
      if (P.dbg && tid == 0) atomicAdd(&g_st_chain_wait, (unsigned long long)(clock64() - tw_));
#endif
    }
    return ST_GO;
  }
  __threadfence();
  const long long t_tick = (!PERSIST && P.dbg) ? clock64() : 0;
  // records of the pending state sum: the blocks of the layout the previous resampling (or the init) ran on
  int nb_pending = 0;
  if (rprev) {
    const StSeg& sp = P.seg[pp * f.C + c];
    nb_pending = obs == 0 ? L.nb : st_make_layout<TS>(sp.goff, sp.nloc, P.bpc).nb;
  }
  StRec r;
  st_local_merge<Real, ST_THREADS>(P, c, L.nb, nb_pending, s_red, r);
  const long long t_merge = (!PERSIST && P.dbg) ? clock64() : 0;
  if (tid == 0) P.counter[c] = 0u;
  if (P.sharded) {
    if (tid == 0) {
      bool fused = false;
      if constexpr (!PERSIST) {   // (the chain-persistent kernel never runs sharded: its text stays free of this branch)
        if (P.peer_seq0) {   // the exchange and the global bookkeeping in this kernel's tail (k_st_merge's work)
          fused = true;
          if (st_peer_allgather(P, c, r, P.peer_seq0 + (unsigned long long)obs))
            st_global(P, c, obs, P.rec_all + c, f.C, P.world, P.rank, L.goff, L.nloc, ST_ALL_ROLES);
          else st_peer_failed(P, c);
        }
      }
      if (!fused) P.rec_local[c] = r;
    }
  }
  else if (lane == 0 && wid < 3) st_global(P, c, obs, &r, 0, 1, 0, L.goff, L.nloc, 1 << wid);   // three roles on three warps
  if constexpr (PERSIST) st_chain_publish(&P.epoch[c], (unsigned int)(obs + 1));
  else if (P.dbg && tid == 0) { P.dbg[0] = t_tick - t_start; P.dbg[1] = t_merge - t_tick; P.dbg[2] = clock64() - t_merge; P.dbg[7] = t_merge - P.dbg[7]; P.dbg[6] = P.dbg[6] - P.dbg[5]; P.dbg[5] = P.dbg[5] - P.dbg[4]; P.dbg[4] = P.dbg[4] - t_tick; }
  return ST_GO;
}
template <typename Model, typename Real, int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS, BSSM_ST_OCC_STEP / THREADS) k_st_step(const __grid_constant__ StreamParams P, int obs) {
  __shared__ StStepSmem<Real, THREADS> sm;
  const FilterDev& f = P.f;
  const int c = blockIdx.x / P.bpc, j = blockIdx.x % P.bpc;
  st_pdl_launch_dependents();
  // independent of the previous launch: parameters, Philox key
  Real par[Model::NPAR];
  Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
  const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  st_step_body<Model, Real, PPT, THREADS, false>(P, obs, c, j, par, key, sm);
}

// sharded runs: global bookkeeping from the gathered records (one thread per filter)
static __global__ void k_st_merge(StreamParams P, int obs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.f.C || !P.f.alive[c]) return;
  const StLayout L = st_layout_in<1>(P, c, obs);
  st_global(P, c, obs, P.rec_all + c, P.f.C, P.world, P.rank, L.goff, L.nloc, ST_ALL_ROLES);
}

// ---- K_B: resampling (scan + closed-form offspring ranges + staged scatter) ----
// ---- multinomial resampling (src/resampling.cpp:5-13) on the streaming engine ------------------------------------------
// n iid uniforms, sorted, are the normalised partial sums of n + 1 iid Exp(1) spacings: U_(i) = (E_0 + .. + E_i) / (E_0 + .. + E_n),
// E_k = -log(U_k), U_k the Philox word of slot k.  The positions arrive in increasing order, so the input-centric expansion of
// k_st_resample serves them like the stratified ones: source j owns the slots [F(c_{j-1}), F(c_j)), F(c) = #{ i : U_(i) <= c } --
// a count found in the staged positions instead of in closed form.  Same law as Rcpp::sample's draws (offspring counts
// ~ Multinomial(n, p)); the ancestors come out sorted (oracle: orc_resample_multinomial_sorted).
// Three small kernels per resampling step: tile sums of the spacings, their scan, the positions.  Fixed reduction trees: the
// positions do not depend on the launch geometry.
constexpr int MN_THREADS = 256, MN_TILE = 1024;
__device__ __forceinline__ void st_mn_spacings(const NoiseKey& key, unsigned int obs, int i0, int n, double* e /*4*/) {
  const uint4x q = noise_quad(key, obs, TAG_RESAMP_U, 0u, (unsigned int)i0 >> 2);
#pragma unroll
  for (int k = 0; k < 4; k++) e[k] = (i0 + k <= n) ? -log(word_to_unit_f64(q.w[k])) : 0.0;   // slots 0 .. n: n + 1 spacings
}
// block-wide sum in a fixed order (warp trees, then the warps one after the other); valid in every thread
__device__ __forceinline__ double st_mn_block_sum(double v, double* s_w /*[MN_THREADS / 32]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) s_w[wid] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < MN_THREADS / 32; w++) t += s_w[w];
  return t;
}
// first row of the observation's copy of the multinomial arrays ([2][C] rows when they are laid out ahead, by parity)
__device__ __forceinline__ int st_mn_row(const StreamParams& P, int obs) { return P.mn_ahead ? (obs & 1) * P.f.C : 0; }
static __global__ void __launch_bounds__(MN_THREADS) k_st_mn_sums(StreamParams P, int obs) {
  __shared__ double s_w[MN_THREADS / 32];
  const FilterDev& f = P.f;
  const int c = blockIdx.x, tile = blockIdx.y;
  if (!P.mn_ahead && (!f.alive[c] || !P.res[(obs & 1) * f.C + c])) return;
  const int n = filt_n(f, c);
  if ((long long)tile * MN_TILE > n) return;
  const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  double e[4];
  st_mn_spacings(key, (unsigned int)obs, tile * MN_TILE + 4 * threadIdx.x, n, e);
  const double t = st_mn_block_sum((e[0] + e[1]) + (e[2] + e[3]), s_w);
  if (threadIdx.x == 0) P.mn_tsum[((size_t)st_mn_row(P, obs) + c) * P.mn_nt + tile] = t;
}
// exclusive prefix of a filter's tile sums (in place) and their total: one block per filter, 256 tiles per round
static __global__ void __launch_bounds__(MN_THREADS) k_st_mn_scan(StreamParams P, int obs) {
  __shared__ double s_w[MN_THREADS / 32];
  const FilterDev& f = P.f;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (!P.mn_ahead && (!f.alive[c] || !P.res[(obs & 1) * f.C + c])) return;
  const int n = filt_n(f, c);
  const int ntile = n / MN_TILE + 1;          // tiles that hold one of the slots 0 .. n
  double* ts = P.mn_tsum + ((size_t)st_mn_row(P, obs) + c) * P.mn_nt;
  double carry = 0.0;
  for (int b0 = 0; b0 < ntile; b0 += MN_THREADS) {
    const int t = b0 + tid;
    const double v = t < ntile ? ts[t] : 0.0;
    const double inc = warp_incl_scan_d(v, lane);
    __syncthreads();
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    double woff = 0.0, tot = 0.0;
#pragma unroll
    for (int w = 0; w < MN_THREADS / 32; w++) { if (w == wid) woff = tot; tot += s_w[w]; }
    if (t < ntile) ts[t] = carry + (woff + (inc - v));
    carry += tot;
  }
  if (tid == 0) P.mn_total[st_mn_row(P, obs) + c] = carry;
}
static __global__ void __launch_bounds__(MN_THREADS) k_st_mn_positions(StreamParams P, int obs) {
  __shared__ double s_w[MN_THREADS / 32];
  const FilterDev& f = P.f;
  const int c = blockIdx.x, tile = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (!P.mn_ahead && (!f.alive[c] || !P.res[(obs & 1) * f.C + c])) return;
  const int n = filt_n(f, c);
  if ((long long)tile * MN_TILE >= n) return;
  const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  const int i0 = tile * MN_TILE + 4 * tid;
  double e[4];
  st_mn_spacings(key, (unsigned int)obs, i0, n, e);
  const double run = (e[0] + e[1]) + (e[2] + e[3]);       // the association k_st_mn_sums used
  const double inc = warp_incl_scan_d(run, lane);
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  double woff = 0.0;
#pragma unroll
  for (int w = 0; w < MN_THREADS / 32; w++) if (w < wid) woff += s_w[w];
  double cum = P.mn_tsum[((size_t)st_mn_row(P, obs) + c) * P.mn_nt + tile] + (woff + (inc - run));
  const double total = P.mn_total[st_mn_row(P, obs) + c];
  double* pos = P.mn_pos + ((size_t)st_mn_row(P, obs) + c) * P.xstride;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    cum += e[k];
    if (i0 + k < n) pos[i0 + k] = cum / total;
  }
}
// #{ i < n : pos[i] <= c }, by one warp: a window around c n (the order statistics stay within a few sqrt(n) slots of the
// diagonal), widened until it brackets the answer, then narrowed 32 ways per round
__device__ __forceinline__ int st_mn_count(const double* __restrict__ pos, int n, double c, int lane) {
  if (!(c > 0.0)) return 0;
  if (c >= 1.0) return n;
  const int g = (int)(c * (double)n);
  int w = (int)(4.0 * sqrt((double)n)) + 64;
  int lo, hi;
  for (;;) {
    lo = max(0, g - w); hi = min(n, g + w);
    const bool ok_lo = lo == 0 || pos[lo - 1] <= c, ok_hi = hi == n || pos[hi] > c;   // answer in [lo, hi]
    if (ok_lo && ok_hi) break;
    w *= 4;
  }
  while (hi - lo > 32) {
    const int stp = (hi - lo + 31) / 32;
    const int idx = lo + (lane + 1) * stp - 1;
    const bool le = idx < hi && pos[idx] <= c;
    const int k = __popc(__ballot_sync(0xffffffffu, le));
    lo += k * stp;
    hi = min(hi, lo + stp);
  }
  const bool le = lo + lane < hi && pos[lo + lane] <= c;
  return lo + __popc(__ballot_sync(0xffffffffu, le));
}

// Same block -> tile ranges and prefetch as k_st_step.  A block's cdf interval comes from the block prefix
// array (bit-identical in the neighbouring blocks); the tile boundaries inside it are this block's own
// running sums of the tile totals it computes itself (clamped into the block's interval).
// PERSIST: see st_step_body; the caller meets the other blocks of the filter after the body (the next observation reads
// slots its neighbours wrote).  Result: ST_LEAVE for a dead filter / a block without tiles, else ST_GO; *resampled tells
// whether the filter resampled at this observation (the same answer in every block of the filter).
template <typename Model, typename Real, int PPT, int THREADS, bool PERSIST>
__device__ __forceinline__ int st_resample_body(const StreamParams& P, int obs, int c, int j, const Real* par, const NoiseKey& key,
                                                StResSmem<Real, PPT, THREADS>& sm, int* resampled, int* nb_out) {
  constexpr int ST_THREADS = THREADS, ST_NW = THREADS / 32;
  constexpr bool F32 = sizeof(Real) == 4;
  constexpr int TS = ST_THREADS * PPT;
  constexpr int CAP = StResSmem<Real, PPT, THREADS>::CAP;
  constexpr int SPT = CAP / ST_THREADS;      // output slots per thread in the expansion
  static_assert(CAP % ST_THREADS == 0 && SPT % 2 == 0, "staging capacity: whole, even number of slots per thread");
  unsigned char* const s_uo = sm.uo;
  unsigned int* const s_u = (unsigned int*)s_uo;
  double* const s_p = (double*)s_uo;
  Real* const s_out = (Real*)s_uo;
  int* const s_mn = sm.mn;
  unsigned int* const s_head = sm.head;
  uint4 (&s_pf)[2][2 * ST_THREADS] = sm.pf;
  double* const s_red = sm.red; double* const s_bs = sm.bs;
  int* const s_wf = sm.wf;
  unsigned int* const s_wh = sm.wh;
  const FilterDev& f = P.f;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // independent of the previous launch: the observation
  const int n = P.n_glob ? P.n_glob : filt_n(f, c);
  const int ot = f.obs_times ? f.obs_times[obs] : obs + 1;
  double yv[4] = {0, 0, 0, 0};
  for (int k = 0; k < f.dy && k < 4; k++) yv[k] = f.y[(size_t)obs * f.dy + k];
  if constexpr (!PERSIST) st_pdl_wait();
  *resampled = 0;
  if (!f.alive[c]) return ST_LEAVE;
  const int pc = obs & 1;
  if (!P.res[pc * f.C + c]) return ST_GO;
  *resampled = 1;
  const StSeg sg = P.seg[pc * f.C + c];
  const StLayout L = st_make_layout<TS>(sg.goff, sg.nloc, P.bpc);
  *nb_out = L.nb;
  if (j >= L.nb) return ST_LEAVE;
  const int lead = L.lead, ntc = L.ntc;
  const int t0 = j * L.tpb, t1 = min(ntc, t0 + L.tpb);
  const Real* xin = (const Real*)P.x1 + (size_t)c * P.xstride + tid * PPT;
  st_prefetch<Real, PPT, ST_THREADS>(s_pf[0], xin + (size_t)t0 * TS);
  const double M = f.M[c], S = f.S[c];
  const double wscale = 1.0 / S;
  const double* pref = P.pref + (size_t)c * (P.bpc + 1);
  const double b_lo = pref[j], b_hi = pref[j + 1];
  Real* xo = (Real*)P.x0 + (size_t)c * P.xstride - (sg.ngoff & ~3LL);   // xo[slot] = storage of global slot
#pragma unroll
  for (int i = 0; i < SPT; i++) s_head[tid * SPT + i] = 0u;     // every thread keeps its own slots of the head array clear
  // cdf numerator interval of this block on the global scale (block edges: the neighbouring block evaluates the same
  // expression on the same prefix value); inside the block the tile edges are running sums of this kernel's own tile totals
  const double blk_lo = st_bound(sg.abase, b_lo, sg.gscale), blk_hi = st_bound(sg.abase, b_hi, sg.gscale);
  double run = 0.0;    // sum of the tile totals of this block so far (identical in every thread)
  double bacc = 0.0;   // this thread's share of the sum of the states written by this block

  for (int tile = t0; tile < t1; tile++) {
  const int pb = (tile - t0) & 1;
  const int sbase = tile * TS + tid * PPT;
  const long long g0 = sg.goff - lead + sbase;
  const int k_lo = (int)max(0LL, min((long long)PPT, sg.goff - g0));
  const int k_hi = (int)max(0LL, min((long long)PPT, sg.goff + sg.nloc - g0));
  const int last_s = min(TS, sg.nloc + lead - tile * TS) - 1;   // in-tile storage index of the last valid particle
  const bool ragged = k_lo > 0 || k_hi < PPT;
  const bool special = ragged || (last_s >= tid * PPT && last_s < (tid + 1) * PPT);
  Real x[PPT], e[PPT];
  st_take<Real, PPT, ST_THREADS>(x, s_pf[pb]);
  if (tile + 1 < t1) st_prefetch<Real, PPT, ST_THREADS>(s_pf[pb ^ 1], xin + (size_t)(tile + 1) * TS);
  Real fs = 0;
  {
    const Real Mr = (Real)M;
    if constexpr (F32) {        // exponent arguments two particles per instruction (packed fp32 pairs)
      const F2 nm2 = f2_make(-Mr, -Mr), l2e = f2_make(1.4426950408889634f, 1.4426950408889634f);
#pragma unroll
      for (int k = 0; k < PPT; k += 2) {
        F2 lw2;
        if constexpr (ModelPacked<Model>::value) lw2 = Model::loglik2(yv, f2_make(x[k], x[k + 1]), par, ot);
        else lw2 = f2_make(Model::template loglik<Real>(yv, &x[k], par, ot), Model::template loglik<Real>(yv, &x[k + 1], par, ot));
        float a0, a1;
        f2_get(f2_mul(f2_add(lw2, nm2), l2e), a0, a1);
        e[k] = Math<float>::ex2_(a0); e[k + 1] = Math<float>::ex2_(a1);
      }
    } else {
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        const Real lw = Model::template loglik<Real>(yv, &x[k], par, ot);
        e[k] = Math<Real>::exp_(lw - Mr);
      }
    }
    if (ragged) {
#pragma unroll
      for (int k = 0; k < PPT; k++) if (k < k_lo || k >= k_hi) e[k] = (Real)0;
    }
#pragma unroll
    for (int k = 0; k < PPT; k++) fs += e[k];
  }
  // tile-local exclusive prefix of the thread sums: fp64 across warps; inside a warp fp32 in the throughput
  // precision (256 particles: error ~1e-4 output slots), fp64 in the parity precision
  double exu, lo_cdf, hi_cdf;
  {
    Real inc = fs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { Real t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_red[wid] = (double)inc;
    __syncthreads();
    double wbase = 0.0, total = 0.0;
#pragma unroll
    for (int w = 0; w < ST_NW; w++) { if (w == wid) wbase = total; total += s_red[w]; }   // fixed order, identical in every thread
    exu = wbase + (double)(inc - fs);
    // tile edges: rank edges from the descriptor, block edges from the prefix array, edges inside the block from the
    // running sum -- clamped into the block's interval, so that rounding can never reach into a neighbour's slots
    double A_lo, A_hi;
    if (tile == 0) A_lo = sg.abase;
    else A_lo = tile == t0 ? blk_lo : fmin(blk_lo + run, blk_hi);
    run += total;
    if (tile == ntc - 1) A_hi = sg.aend;
    else A_hi = tile == t1 - 1 ? blk_hi : fmin(blk_lo + run, blk_hi);
    if (F32) { lo_cdf = A_lo * wscale; hi_cdf = A_hi * wscale; }      // reciprocal: no fp64 division chain
    else { lo_cdf = A_lo / S; hi_cdf = A_hi / S; }
  }
  const bool tail = sg.last && tile == ntc - 1;
  if (tail) hi_cdf = 2.0;

  SlotCounter sc;
  sc.key = key; sc.obs = (unsigned int)obs; sc.fn = P.resample_fn; sc.n = n; sc.w_sys = 0u;
  sc.s_u = s_u; sc.u_cap = CAP;
  {
    const double t0 = lo_cdf * (double)n;
    const int i0 = t0 >= (double)n ? n : (int)t0;
    sc.u_base = max(0, (i0 & ~3) - 4);
  }
  const bool MN = P.resample_fn == 2;
  const double* const mpos = P.mn_pos + ((size_t)(P.mn_ahead ? (obs & 1) * f.C : 0) + c) * P.xstride;   // multinomial: the sorted positions of the output slots
  int o_lo, o_hi, mn_cnt = 0;
  if (MN) {
    // output range of the tile: counts of positions below its cdf interval's ends (warp 0; the neighbouring tile / block
    // counts against the same value), then the positions of that range into shared memory
    if (wid == 0) {
      const int a = st_mn_count(mpos, n, lo_cdf, lane);
      const int z = tail ? n : st_mn_count(mpos, n, hi_cdf, lane);
      if (lane == 0) { s_mn[0] = a; s_mn[1] = z; }
    }
    __syncthreads();
    o_lo = s_mn[0]; o_hi = max(s_mn[1], o_lo);
    mn_cnt = min(o_hi - o_lo, CAP);
    for (int i = tid; i < mn_cnt; i += ST_THREADS) s_p[i] = mpos[o_lo + i];
    __syncthreads();
  } else {
    if (P.resample_fn == 1) {
      uint4x q0 = noise_quad(key, (unsigned int)obs, TAG_RESAMP_U, 0u, 0u);
      sc.w_sys = q0.w[0];
    } else {
      // the window ends with the quad after the slot the tile's upper cdf edge falls into (an fp32 position may round one slot
      // past the edge; what it finds there is clamped to o_hi), not with the buffer: a Philox call per four slots is the
      // largest single item of this kernel
      const double th = hi_cdf * (double)n;
      const int i_hi = th >= (double)n ? n : (int)th;
      const int q_end = min(min((n + 3) >> 2, (sc.u_base + CAP) >> 2), (i_hi >> 2) + 2);
      sc.u_cap = max(0, min(CAP, 4 * q_end - sc.u_base));
      for (int qd = (sc.u_base >> 2) + tid; qd < q_end; qd += ST_THREADS) {
        uint4x uq = philox4x32_10_rk(P.rk0, (unsigned int)qd, (unsigned int)obs, key.stream, TAG_RESAMP_U, key.k1);
        *(uint4*)&s_u[4 * qd - sc.u_base] = make_uint4(uq.w[0], uq.w[1], uq.w[2], uq.w[3]);
      }
    }
    __syncthreads();
    o_lo = sc.count_le(lo_cdf);
    o_hi = tail ? n : sc.count_le(hi_cdf);
  }
  int F[PPT];
  {
    int fmax = o_lo;
    if (MN) {
      // F(c) = o_lo + #{ staged positions <= c }: a binary search for the thread's first particle, a short walk for the next ones
      // (the positions grow with k); beyond the staged window (more than CAP offspring in the tile) a search in global memory
      double acc = exu;
      int j = 0;
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        acc += (double)e[k];
        int v = o_lo;
        if (k >= k_lo && k < k_hi) {
          const double ck = lo_cdf + acc * wscale;
          if (k == k_lo) {
            int lo = 0, hi = mn_cnt;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_p[mid] <= ck) lo = mid + 1; else hi = mid; }
            j = lo;
          } else {
            while (j < mn_cnt && s_p[j] <= ck) j++;
          }
          v = o_lo + j;
          if (j == mn_cnt && o_hi - o_lo > mn_cnt) {
            int lo = o_lo + mn_cnt, hi = o_hi;
            while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (mpos[mid] <= ck) lo = mid + 1; else hi = mid; }
            v = lo;
          }
          if (tid * PPT + k == last_s) v = o_hi;
          v = min(max(v, o_lo), o_hi);
        }
        fmax = max(fmax, v);
        F[k] = fmax;
      }
    } else if (F32) {
      const double T0 = (lo_cdf + exu * wscale) * (double)n;
      const double T0c = T0 < (double)n ? T0 : (double)n;
      const int I0 = (int)T0c;
      const float f0 = (float)(T0c - (double)I0);
      const float wsn = (float)(wscale * (double)n);
      float accf = 0.f;
      // One test per thread instead of three per particle: the positions grow with k, so if the LAST one is inside
      // the staged window of uniforms, below n and within the range of the fp32 floor trick, all of them are.
      const float tf_last = fmaf((float)fs, wsn, f0);
      const int rel0 = I0 - sc.u_base;
      const bool fast = P.resample_fn != 1 && tf_last < 4194304.0f && rel0 >= 0 && rel0 + (int)tf_last + 1 < sc.u_cap &&
                        I0 + (int)tf_last + 1 < n;
      if (fast) {
        const unsigned int* su = s_u + rel0;
        // two particles per instruction (packed fp32 pairs) wherever the arithmetic is the same on both
        const F2 wsn2 = f2_make(wsn, wsn), f02 = f2_make(f0, f0), mh2 = f2_make(-0.5f, -0.5f), mg2 = f2_make(12582912.0f, 12582912.0f),
                 nmg2 = f2_make(-12582912.0f, -12582912.0f), one2 = f2_make(1.0f, 1.0f), m12 = f2_make(-1.0f, -1.0f);
#pragma unroll
        for (int k = 0; k < PPT; k += 2) {
          const float acc0 = accf + (float)e[k];
          accf = acc0 + (float)e[k + 1];
          const F2 tf = f2_fma(f2_make(acc0, accf), wsn2, f02);
          const F2 r = f2_add(f2_add(tf, mh2), mg2);                         // floor and fraction without conversion instructions
          const F2 g = f2_add(f2_fma(f2_add(r, nmg2), m12, tf), one2);         // 1 + fraction, in [1, 2]
          float r0, r1, g0, g1;
          f2_get(r, r0, r1); f2_get(g, g0, g1);
          const int ii0 = __float_as_int(r0) - 0x4B400000, ii1 = __float_as_int(r1) - 0x4B400000;
          // (i + U_i) <= t  <=>  U_i <= frac, on the 23 leading bits of the word: 1.U_i < 1 + frac as floats (g = 2: always)
          const float u0 = __uint_as_float(0x3F800000u | (su[ii0] >> 9)), u1 = __uint_as_float(0x3F800000u | (su[ii1] >> 9));
          // the positions grow with k (accf sums non-negative terms), so v does: no running maximum; the lower clamp is
          // the max with prevF below
          F[k] = min(I0 + ii0 + (u0 < g0 ? 1 : 0), o_hi);
          F[k + 1] = min(I0 + ii1 + (u1 < g1 ? 1 : 0), o_hi);
        }
        fmax = F[PPT - 1];
      } else {
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          accf += (float)e[k];
          const float tf = fmaf(accf, wsn, f0);
          const float r = (tf - 0.5f) + 12582912.0f;
          const int ii = __float_as_int(r) - 0x4B400000;
          const float frac = tf - (r - 12582912.0f);
          const float g = frac + 1.0f;
          const unsigned int fbits = ((unsigned int)__float_as_int(g) & 0x7FFFFFu) << 9;
          const int i = I0 + ii;
          int v;
          if (tf >= 4194304.0f) v = sc.count_le(lo_cdf + (exu + (double)accf) * wscale);   // beyond the fp32 floor trick (degenerate weights)
          else if (i >= n) v = n;
          else v = i + ((sc.word_of(i) < fbits || g >= 2.0f) ? 1 : 0);
          v = min(max(v, o_lo), o_hi);
          fmax = max(fmax, v);
          F[k] = fmax;
        }
      }
      if (special) {   // the thread with padding lanes and / or the last valid particle of the tile (which takes what is left)
        fmax = o_lo;
#pragma unroll
        for (int k = 0; k < PPT; k++) {
          int v = F[k];
          if (tid * PPT + k == last_s) v = o_hi;
          if (k < k_lo || k >= k_hi) v = o_lo;
          fmax = max(fmax, v);
          F[k] = fmax;
        }
      }
    } else {
      double acc = exu;
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        acc += (double)e[k];
        int v = o_lo;
        if (k >= k_lo && k < k_hi) {
          v = sc.count_le(lo_cdf + acc * wscale);
          if (tid * PPT + k == last_s) v = o_hi;
          v = min(max(v, o_lo), o_hi);
        }
        fmax = max(fmax, v);
        F[k] = fmax;
      }
    }
  }
  int prevF;
  {
    int inc = max(F[PPT - 1], o_lo);   // (the fast path above leaves the lower clamp to this scan)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, t); }
    if (lane == 31) s_wf[wid] = inc;
    prevF = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prevF = o_lo;
    __syncthreads();
    int wv = lane < ST_NW ? s_wf[lane] : o_lo;
    int winc = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc = max(winc, t); }
    int wprev = __shfl_sync(0xffffffffu, winc, (wid + 31) & 31);
    if (wid == 0) wprev = o_lo;
    prevF = max(prevF, wprev);
#pragma unroll
    for (int k = 0; k < PPT; k++) F[k] = max(F[k], prevF);
  }
  // Expansion, output-centric (chunks of CAP slots): every source with offspring in the chunk marks the first of
  // its slots with (source index, address of its x in the prefetch buffer); a running maximum over the slots --
  // source indices grow with the slot -- tells every slot its source.  O(1) per source and per slot, no loop
  // over the offspring of a source, no special case for heavy sources.  The chosen x are staged and leave the
  // SM as coalesced vector stores.
  Real sumx = 0;
  const int o_base = o_lo & ~3;
  constexpr int EPH = 16 / (int)sizeof(Real);                          // elements per 16-byte half of the prefetch layout
  const Real* s_x = (const Real*)s_pf[pb];
  for (int c0 = o_base; c0 < o_hi; c0 += CAP) {
    const int c1 = min(o_hi, c0 + CAP);
    int lo_k = prevF;
    const unsigned int key0 = ((unsigned int)(tid * PPT) << 16) | (unsigned int)(tid * EPH);   // key of source k = key0 + a constant
    if (o_hi - o_base <= CAP) {          // the usual case: all the offspring of the tile in one chunk (block-uniform)
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        const int hi_k = F[k];           // F is non-decreasing and starts at prevF
        // branch-free: a source without offspring writes its mark into the thread's own spare word behind the head array
        unsigned int* const hp = (hi_k > lo_k) ? &s_head[lo_k - o_base] : &s_head[CAP + tid];
        *hp = key0 + (((unsigned int)k << 16) | (unsigned int)((k / EPH) * ST_THREADS * EPH + (k % EPH)));
        if (hi_k > lo_k) {
          if constexpr (F32) sumx = fmaf(__int_as_float(0x4B000000 | (hi_k - lo_k)) - 8388608.0f, x[k], sumx);   // count as a float, exact below 2^23
          else sumx += (Real)(hi_k - lo_k) * x[k];
        }
        lo_k = hi_k;
      }
    } else {
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        const int hi_k = F[k];
        const int a = max(lo_k, c0);
        if (min(hi_k, c1) > a) s_head[a - c0] = key0 + (((unsigned int)k << 16) | (unsigned int)((k / EPH) * ST_THREADS * EPH + (k % EPH)));
        if (c0 == o_base && hi_k > lo_k) sumx += (Real)(hi_k - lo_k) * x[k];
        lo_k = max(lo_k, hi_k);
      }
    }
    __syncthreads();
    // running maximum over the slots: thread-local, then across the warp, then across the warps
    unsigned int h[SPT];
#pragma unroll
    for (int i = 0; i < SPT; i += 2) { const uint2 v2 = *(const uint2*)&s_head[tid * SPT + i]; h[i] = v2.x; h[i + 1] = v2.y; }
#pragma unroll
    for (int i = 0; i < SPT; i += 2) *(uint2*)&s_head[tid * SPT + i] = make_uint2(0u, 0u);   // clear for the next chunk / tile
#pragma unroll
    for (int i = 1; i < SPT; i++) h[i] = max(h[i], h[i - 1]);
    unsigned int inc = h[SPT - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, t); }
    if (lane == 31) s_wh[wid] = inc;
    unsigned int carry = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) carry = 0u;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < ST_NW - 1; w++) carry = max(carry, (w < wid) ? s_wh[w] : 0u);
    Real val[SPT];
#pragma unroll
    for (int i = 0; i < SPT; i++) val[i] = s_x[max(h[i], carry) & 0xFFFFu];
    if constexpr (F32) {
#pragma unroll
      for (int i = 0; i < SPT; i += 2) *(float2*)&s_out[tid * SPT + i] = make_float2(val[i], val[i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < SPT; i++) s_out[tid * SPT + i] = val[i];
    }
    __syncthreads();
    const int first = max(c0, o_lo), last = c1;
    if (F32) {
      for (int o = c0 + 4 * tid; o < last; o += 4 * ST_THREADS) {
        const float4 v = *(const float4*)&s_out[o - c0];
        if (o >= first && o + 3 < last) *(float4*)&xo[o] = v;
        else {
          if (o >= first && o < last) xo[o] = v.x;
          if (o + 1 >= first && o + 1 < last) xo[o + 1] = v.y;
          if (o + 2 >= first && o + 2 < last) xo[o + 2] = v.z;
          if (o + 3 >= first && o + 3 < last) xo[o + 3] = v.w;
        }
      }
    } else {
      for (int o = first + tid; o < last; o += ST_THREADS) xo[o] = s_out[o - c0];
    }
    __syncthreads();
  }
  bacc += (double)sumx;   // this thread's share of the sum of the chosen states; reduced once, after the last tile
  }  // tiles
  // sum of the states written by this block (state estimate after resampling, merged by the next observation's k_st_step)
  {
    const double v = warp_sum_d(bacc);
    if (lane == 0) s_bs[wid] = v;
    __syncthreads();
    bacc = 0.0;
    if (tid == 0) {
      for (int w = 0; w < ST_NW; w++) bacc += s_bs[w];
    }
  }
  if (tid == 0) P.bsum[(size_t)c * P.bpc + j] = bacc;
  return ST_GO;
}
template <typename Model, typename Real, int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS, BSSM_ST_OCC_RES / THREADS) k_st_resample(const __grid_constant__ StreamParams P, int obs) {
  __shared__ StResSmem<Real, PPT, THREADS> sm;
  const FilterDev& f = P.f;
  const int c = blockIdx.x / P.bpc, j = blockIdx.x % P.bpc;
  st_pdl_launch_dependents();
  // independent of the previous launch: parameters, Philox key
  Real par[Model::NPAR];
  Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
  const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  int resampled, nb;
  st_resample_body<Model, Real, PPT, THREADS, false>(P, obs, c, j, par, key, sm, &resampled, &nb);
}

// ---- chain-persistent kernel: every observation of a batch of filters in ONE cooperative launch ----
// Block (c, j) keeps its tile range of filter c for the whole run and alternates the two bodies above.  The blocks of a filter
// meet twice per observation -- after the merge of the block records (the merging block publishes the bookkeeping) and after
// a resampling (the next step reads slots the neighbouring blocks wrote) -- through two words per filter; filters never wait
// for each other, so the blocks an SM holds drift apart and fill each other's waits.  Launched cooperatively (the blocks of
// a filter spin on each other: all of them must be resident); serves batches, where two launches per observation leave the
// chip a quarter idle (ramp, drain, every block in the same phase).
template <typename Model, typename Real, int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS, BSSM_ST_OCC_RES / THREADS) k_st_chain(const __grid_constant__ StreamParams P, int c_base, int T) {
  union Smem { StStepSmem<Real, THREADS> a; StResSmem<Real, PPT, THREADS> b; };
  __shared__ Smem sm;
  const FilterDev& f = P.f;
  const int c = c_base + blockIdx.x / P.bpc, j = blockIdx.x % P.bpc;
  Real par[Model::NPAR];
  Model::template prepare<Real>(f.theta + (size_t)c * f.theta_stride, par);
  const NoiseKey key = make_key(f.seed, f.run_id[c], f.stream[c]);
  unsigned int arrivals = 0u;
#ifdef BSSM_ST_CHAIN_TIMING   // diagnostics build: where a block's cycles go (thread 0; step body incl. its wait, resample body, second wait)
  long long tc[3] = {0, 0, 0};
#define ST_TC(i, ...) { const long long t0_ = clock64(); __VA_ARGS__; tc[i] += clock64() - t0_; }
#else
#define ST_TC(i, ...) { __VA_ARGS__; }
#endif
  for (int obs = 0; obs < T; obs++) {
    int leave = 0;
    ST_TC(0, leave = st_step_body<Model, Real, PPT, THREADS, true>(P, obs, c, j, par, key, sm.a) == ST_LEAVE);
    if (leave) return;
    __syncthreads();                       // the two bodies overlay their shared memory
    int resampled = 0, nb = 1;
    ST_TC(1, leave = st_resample_body<Model, Real, PPT, THREADS, true>(P, obs, c, j, par, key, sm.b, &resampled, &nb) == ST_LEAVE);
    if (leave) return;
    if (resampled) {
      arrivals += (unsigned int)nb;
      ST_TC(2, if (nb > 1) st_chain_arrive_wait(&P.bar2[c], arrivals); else __syncthreads());
    }
  }
#ifdef BSSM_ST_CHAIN_TIMING
  if (P.dbg && threadIdx.x == 0) {
    for (int i = 0; i < 3; i++) atomicAdd((unsigned long long*)&P.dbg[i], (unsigned long long)tc[i]);
    atomicAdd((unsigned long long*)&P.dbg[3], 1ull);
    P.dbg[4] = (long long)atomicAdd(&g_st_chain_wait, 0ull);   // the last block to finish leaves the total
  }
#endif
#undef ST_TC
}

// ---- flush: the state estimate of a final resampling (or of the initial particles when T = 0) ----
template <int TS>
static __global__ void __launch_bounds__(256) k_st_flush(StreamParams P, int obs /* = T */) {
  constexpr int ST_THREADS = 256, ST_NW = 8;
  __shared__ double s_red[ST_NW];
  const FilterDev& f = P.f;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (!f.alive[c]) return;
  const int pp = (obs + 1) & 1;
  const int rprev = P.res[pp * f.C + c];
  const StSeg& sp = P.seg[pp * f.C + c];
  // the blocks of the layout the last resampling (or the init) ran on
  const int nbp = st_make_layout<TS>(sp.goff, sp.nloc, P.bpc).nb;
  double lp = 0.0;
  if (rprev) for (int j = tid; j < nbp; j += ST_THREADS) lp += P.bsum[(size_t)c * P.bpc + j];
  lp = warp_sum_d(lp);
  if (lane == 0) s_red[wid] = lp;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < ST_NW; w++) t += s_red[w];
    const int n = P.n_glob ? P.n_glob : filt_n(f, c);
    if (P.sharded) {
      StRec r; r.m = 0; r.s = 0; r.q = 0; r.sx = 0; r.pend = t; r.nan = rprev ? 1.0 : 0.0; r.pad0 = r.pad1 = 0;
      if (P.peer_seq0) {   // k_st_flush_merge's work, after the same fused exchange as in k_st_step
        if (!st_peer_allgather(P, c, r, P.peer_seq0 + (unsigned long long)obs)) st_peer_failed(P, c);
        else if (rprev) {
          double tt = 0.0;
          for (int g = 0; g < P.world; g++) tt += P.rec_all[(size_t)g * f.C + c].pend;
          if (obs == 0) f.ess[(size_t)c * (f.T + 1)] = (double)P.n_glob;
          f.state_est[(size_t)c * (f.T + 1) + obs] = tt / (double)P.n_glob;
        }
      } else P.rec_local[c] = r;
    }
    else if (rprev) {
      if (obs == 0) f.ess[(size_t)c * (f.T + 1)] = (double)n;
      f.state_est[(size_t)c * (f.T + 1) + obs] = t / (double)n;
    }
  }
}
static __global__ void k_st_flush_merge(StreamParams P, int obs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const FilterDev& f = P.f;
  if (c >= f.C || !f.alive[c]) return;
  if (!P.res[((obs + 1) & 1) * f.C + c]) return;
  double t = 0.0;
  for (int g = 0; g < P.world; g++) t += P.rec_all[(size_t)g * f.C + c].pend;
  if (obs == 0) f.ess[(size_t)c * (f.T + 1)] = (double)P.n_glob;
  f.state_est[(size_t)c * (f.T + 1) + obs] = t / (double)P.n_glob;
}

}  // namespace bssm
