// bssm_resample.cuh -- resampling kernels, batched over independent
// weight vectors ("segments": filters / chains).  Replaces src/resampling.cpp:5-66.
//
//   cdf:    k_tile_sums -> k_tile_scan -> [k_chain -> k_tile_exact]      (fp64 always)
//   search: k_search (stratified / systematic / multinomial), first j with cdf[j] >= pos,
//           clamp n-1 (src/resampling.cpp:32-37,59-63)
//
// "exact" mode reproduces the reference's sequential double sum / cumsum bit for bit
// (bssm_exact.cuh); "fast" mode stops after the ordinary parallel scan.
#pragma once
#include "bssm_common.cuh"
#include "bssm_exact.cuh"

namespace bssm {

constexpr int RS_THREADS = 256;
constexpr int RS_IPT = 4;
constexpr int RS_TILE = RS_THREADS * RS_IPT;  // 1024 elements per tile

struct TileRec {
  ParFn fa, fb;   // maps for binade be / alt
  int be, alt;    // biased exponents; alt < 0: none
  int regular;    // 0: the chain walks this tile serially
  int pad;
};

// ---- value sources -------------------------------------------------------------------------
// v(seg, i) >= 0 in double.  `total` (when used) is the exact sequential sum from a first pass.
struct SrcPlain {  // raw weights [seg][stride]
  const double* w; size_t stride;
  __device__ __forceinline__ double operator()(int seg, int i) const { return w[(size_t)seg * stride + i]; }
};
struct SrcPlainNorm {  // w / total   (src/resampling.cpp:24,51)
  const double* w; size_t stride; const double* total;
  __device__ __forceinline__ double operator()(int seg, int i) const { return w[(size_t)seg * stride + i] / total[seg]; }
};
// filter weights from stored log-weights: w = exp(lw - M) / S  (R/particle_filter_core.R:204-207)
template <typename Real> struct SrcLogW {
  const Real* lw; size_t stride; const double* M; const double* S;
  __device__ __forceinline__ double operator()(int seg, int i) const {
    return exp((double)lw[(size_t)seg * stride + i] - M[seg]) / S[seg];
  }
};
template <typename Real> struct SrcLogWNorm {  // (exp(lw - M) / S) / total
  const Real* lw; size_t stride; const double* M; const double* S; const double* total;
  __device__ __forceinline__ double operator()(int seg, int i) const {
    return (exp((double)lw[(size_t)seg * stride + i] - M[seg]) / S[seg]) / total[seg];
  }
};

// ---- block helpers -------------------------------------------------------------------------
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ double block_sum_bcast(double v, double* sm /* >= 33 */) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) sm[32] = t;
  }
  __syncthreads();
  return sm[32];
}
__device__ __forceinline__ i64 shfl_up_i64(i64 v, int o) {
  int lo = __shfl_up_sync(0xffffffffu, (int)(v & 0xffffffffll), o);
  int hi = __shfl_up_sync(0xffffffffu, (int)(v >> 32), o);
  return ((i64)hi << 32) | (unsigned int)lo;
}

// per-segment enable flag: enable == nullptr -> always on
__device__ __forceinline__ bool seg_on(const int* enable, int seg) { return enable == nullptr || enable[seg] != 0; }

// ---- K4a: approximate tile sums (+ validation on the raw weights) -----------------------------
// status[seg]: BSSM_ERR_* by atomicMax (NaN=3 > negative=1)
template <typename Src>
__global__ void __launch_bounds__(RS_THREADS) k_tile_sums(Src src, int n, const int* __restrict__ n_per_seg,
                                                         int ntiles, double* __restrict__ part,
                                                         int* __restrict__ status, int validate,
                                                         const int* __restrict__ enable) {
  __shared__ double sm[34];
  int seg = blockIdx.x, tile = blockIdx.y;   // segment in grid.x (no 65535 limit)
  if (!seg_on(enable, seg)) return;
  int nn = n_per_seg ? n_per_seg[seg] : n;
  int base = tile * RS_TILE + threadIdx.x * RS_IPT;
  double s = 0.0;
  int bad = 0;
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    int i = base + k;
    if (i < nn) {
      double v = src(seg, i);
      if (validate) { if (v != v) bad = 3; else if (v < 0 && bad < 1) bad = 1; }
      s += v;
    }
  }
  double tot = block_sum_bcast(s, sm);
  if (threadIdx.x == 0) part[(size_t)seg * ntiles + tile] = tot;
  if (validate && bad) atomicMax(&status[seg], bad);
}

// ---- exclusive prefix of a segment's tile totals (large inputs) --------------------------------
// Below RS_PREFIX_TILES tiles k_tile_scan sums the preceding totals itself (<= 8 loads per thread, no launch); beyond, that
// sum would be O(tiles^2) loads per scan, so one block per segment scans the totals once (256 per round, fixed order).
constexpr int RS_PREFIX_TILES = 2048;
static __global__ void __launch_bounds__(256) k_tile_prefix(const double* __restrict__ part, int n, const int* __restrict__ n_per_seg,
                                                            int ntiles, double* __restrict__ pref, const int* __restrict__ enable) {
  __shared__ double wsum[8];
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (!seg_on(enable, seg)) return;
  const int nn = n_per_seg ? n_per_seg[seg] : n;
  const int nt = (nn + RS_TILE - 1) / RS_TILE;
  double carry = 0.0;
  for (int b0 = 0; b0 < nt; b0 += 256) {
    const int t = b0 + tid;
    const double v = t < nt ? part[(size_t)seg * ntiles + t] : 0.0;
    const double inc = warp_incl_scan(v, lane);
    __syncthreads();
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    double woff = 0.0, tot = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) { if (w == wid) woff = tot; tot += wsum[w]; }
    if (t < nt) pref[(size_t)seg * ntiles + t] = carry + (woff + (inc - v));
    carry += tot;
  }
}

// ---- K4b: approximate inclusive scan per tile, binade classification, tile maps ---------------
// mode 0 (fast): write the approximate cdf to `cdf`.  mode 1 (exact): write TileRec only.
template <typename Src>
__global__ void __launch_bounds__(RS_THREADS) k_tile_scan(Src src, int n, const int* __restrict__ n_per_seg,
                                                         int ntiles, const double* __restrict__ part,
                                                         double* __restrict__ cdf, size_t cdf_stride,
                                                         TileRec* __restrict__ rec, int exact,
                                                         const int* __restrict__ enable, const double* __restrict__ pref) {
  __shared__ double sm[34];
  __shared__ double wsum[RS_THREADS / 32];
  __shared__ int s_irreg;
  __shared__ i64 fsm[4][RS_THREADS / 32];
  int seg = blockIdx.x, tile = blockIdx.y;   // segment in grid.x (no 65535 limit)
  if (!seg_on(enable, seg)) return;
  int nn = n_per_seg ? n_per_seg[seg] : n;
  if (tile * RS_TILE >= nn) return;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // exclusive tile prefix: fixed-order sum of the preceding partials, or (large inputs) the scanned totals of k_tile_prefix
  if (threadIdx.x == 0) s_irreg = 0;
  double prefix;
  if (pref) { prefix = pref[(size_t)seg * ntiles + tile]; __syncthreads(); }
  else {
    double pre = 0.0;
    for (int t = threadIdx.x; t < tile; t += RS_THREADS) pre += part[(size_t)seg * ntiles + t];
    prefix = block_sum_bcast(pre, sm);
  }
  int base = tile * RS_TILE + threadIdx.x * RS_IPT;
  double v[RS_IPT], loc[RS_IPT];
  double run = 0.0;
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    int i = base + k;
    v[k] = (i < nn) ? src(seg, i) : 0.0;
    run += v[k];
    loc[k] = run;
  }
  double inc = warp_incl_scan(run, lane);
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  double woff = 0.0;
  for (int w = 0; w < wid; w++) woff += wsum[w];
  double excl = prefix + (woff + (inc - run));  // approximate sum before this thread's first element
  double c_prev = excl;
  if (!exact) {
#pragma unroll
    for (int k = 0; k < RS_IPT; k++) {
      int i = base + k;
      if (i < nn) cdf[(size_t)seg * cdf_stride + i] = excl + loc[k];
    }
    return;
  }
  // binade of the approximate running sum before / after each element
  int be_tile = biased_exp(prefix);
  int irreg = (be_tile < 2 || be_tile > 2044) ? 1 : 0;
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    int i = base + k;
    if (i < nn) {
      double c = excl + loc[k];
      if (biased_exp(c_prev) != be_tile || biased_exp(c) != be_tile) irreg = 1;
      c_prev = c;
    }
  }
  if (irreg) s_irreg = 1;  // benign race: all writers store 1
  // last approximate value of the tile (for the near-power-of-two test)
  __shared__ double s_last;
  int last_i = min(nn, (tile + 1) * RS_TILE) - 1;
  if (last_i >= base && last_i < base + RS_IPT) s_last = excl + loc[last_i - base];
  __syncthreads();
  TileRec r;
  r.regular = s_irreg ? 0 : 1;
  r.be = be_tile; r.alt = -1; r.pad = 0;
  r.fa = parfn_identity(); r.fb = parfn_identity();
  if (r.regular) {
    const u64 MASK = (1ull << 52) - 1, NEAR = 1ull << 26;
    u64 f0 = dbits(prefix) & MASK, f1 = dbits(s_last) & MASK;
    if (f1 >= MASK + 1 - NEAR || f0 >= MASK + 1 - NEAR) r.alt = be_tile + 1;
    else if (f0 < NEAR || f1 < NEAR) r.alt = be_tile - 1;
    // block reduction of the maps in element order
    ParFn fa = parfn_identity(), fb = parfn_identity();
#pragma unroll
    for (int k = 0; k < RS_IPT; k++) {
      if (base + k < nn) {
        fa = parfn_compose(fa, parfn_element(v[k], be_tile));
        if (r.alt >= 0) fb = parfn_compose(fb, parfn_element(v[k], r.alt));
      }
    }
    // warp inclusive scan (ordered composition), lane 31 holds the warp total
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      ParFn ga, gb;
      ga.a0 = shfl_up_i64(fa.a0, o); ga.a1 = shfl_up_i64(fa.a1, o);
      gb.a0 = shfl_up_i64(fb.a0, o); gb.a1 = shfl_up_i64(fb.a1, o);
      if (lane >= o) { fa = parfn_compose(ga, fa); fb = parfn_compose(gb, fb); }
    }
    if (lane == 31) { fsm[0][wid] = fa.a0; fsm[1][wid] = fa.a1; fsm[2][wid] = fb.a0; fsm[3][wid] = fb.a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      ParFn ta = parfn_identity(), tb = parfn_identity();
      for (int w = 0; w < RS_THREADS / 32; w++) {
        ParFn xa, xb; xa.a0 = fsm[0][w]; xa.a1 = fsm[1][w]; xb.a0 = fsm[2][w]; xb.a1 = fsm[3][w];
        ta = parfn_compose(ta, xa); tb = parfn_compose(tb, xb);
      }
      r.fa = ta; r.fb = tb;
    }
  }
  if (threadIdx.x == 0) rec[(size_t)seg * ntiles + tile] = r;
}

// ---- K4c: serial chain over the tiles of a segment (one warp per segment) ----------------------
// Produces the exact running sum at every tile start, the binade each regular tile really
// used (used_be, -1 = walked serially here) and the exact total.  Tiles walked serially get
// their cdf written here when cdf != nullptr.  Lanes stage tile records / values through
// shared memory; lane 0 carries the (inherently sequential) floating-point state.
template <typename Src>
__global__ void __launch_bounds__(32) k_chain(Src src, int n, const int* __restrict__ n_per_seg, int ntiles,
                                              const TileRec* __restrict__ rec, double* __restrict__ cstart,
                                              int* __restrict__ used_be, double* __restrict__ total_out,
                                              double* __restrict__ cdf, size_t cdf_stride,
                                              long long* __restrict__ n_serial, const int* __restrict__ enable) {
  __shared__ TileRec srec[32];
  __shared__ double s_cstart[32];
  __shared__ int s_use[32];
  __shared__ double s_c;
  int seg = blockIdx.x, lane = threadIdx.x;
  if (!seg_on(enable, seg)) return;
  int nn = n_per_seg ? n_per_seg[seg] : n;
  int nt = (nn + RS_TILE - 1) / RS_TILE;
  double c = 0.0;  // meaningful in lane 0
  long long serial = 0;
  for (int t0 = 0; t0 < nt; t0 += 32) {
    int cnt = min(32, nt - t0);
    if (lane < cnt) srec[lane] = rec[(size_t)seg * ntiles + t0 + lane];
    __syncwarp();
    // Fast path for the bulk of the cdf: every tile of the batch regular with a map for the binade the running sum is in, and the
    // sum staying in that binade at every tile start.  The maps C -> C + a[C & 1] compose associatively (parfn_compose), so
    // a warp scan gives every tile its start at once -- the very values the walk below would reach one tile after the other
    // (same map choice, same checks); any doubt falls through to that walk.
    bool batch_done = false;
    {
      const double c0 = __shfl_sync(0xffffffffu, c, 0);
      const int bc = biased_exp(c0);
      bool ok = true;
      ParFn G = parfn_identity();
      if (lane < cnt) {
        const TileRec& r = srec[lane];
        if (r.regular && r.be == bc) G = r.fa;
        else if (r.regular && r.alt >= 0 && r.alt == bc) G = r.fb;
        else ok = false;
      }
      if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {          // inclusive scan: G_l = F_0 then F_1 ... then F_l
          ParFn P; P.a0 = shfl_up_i64(G.a0, o); P.a1 = shfl_up_i64(G.a1, o);
          if (lane >= o) G = parfn_compose(P, G);
        }
        const i64 C0 = to_units(c0, bc);
        ParFn Gp; Gp.a0 = shfl_up_i64(G.a0, 1); Gp.a1 = shfl_up_i64(G.a1, 1);
        const i64 Cs = lane == 0 ? C0 : parfn_apply(Gp, C0);     // running sum, in units, before tile `lane` ...
        const i64 Ce = parfn_apply(G, C0);                       // ... and after it
        const bool valid = C0 >= 0 && (lane >= cnt || (units_ok_start(Cs) && units_ok_end(Ce)));
        if (__all_sync(0xffffffffu, valid)) {
          if (lane < cnt) { s_cstart[lane] = from_units(Cs, bc); s_use[lane] = bc; }
          const int lo32 = __shfl_sync(0xffffffffu, (int)(Ce & 0xffffffffll), cnt - 1), hi32 = __shfl_sync(0xffffffffu, (int)(Ce >> 32), cnt - 1);
          c = from_units(((i64)hi32 << 32) | (unsigned int)lo32, bc);
          batch_done = true;
        }
      }
    }
    for (int j = 0; j < cnt && !batch_done; j++) {
      int use = -1;
      if (lane == 0) {
        const TileRec& r = srec[j];
        s_cstart[j] = c;
        if (r.regular) {
          int bc = biased_exp(c);
          if (bc == r.be) use = r.be;
          else if (r.alt >= 0 && bc == r.alt) use = r.alt;
        }
        if (use >= 0) {
          i64 C = to_units(c, use);
          i64 C2 = parfn_apply(use == r.be ? r.fa : r.fb, C);
          if (units_ok_start(C) && units_ok_end(C2)) c = from_units(C2, use); else use = -1;
        }
        s_use[j] = use;
      }
      use = __shfl_sync(0xffffffffu, use, 0);
      if (use < 0) {  // walk the tile in reference order
        int lo = (t0 + j) * RS_TILE, hi = min(nn, lo + RS_TILE);
        serial += hi - lo;
        // The reference's order, one addition after the other -- but in EVERY lane at once: each lane receives the 32 values by
        // shuffle and runs the same chain in registers (8 cycles per addition: the fp64 latency), keeping the partial sum that
        // belongs to its own element.  At N = 2^18 the ~9 tiles in which the running sum changes binade are most of an
        // exact-mode filter's time (k_chain ~200 us, 4 per APF observation).  Measured: a whole block evaluating the tile's values
        // at once cut the instructions 4x and the time not at all -- what remains is the one-thread map chain over the 256 tiles
        // (~100 dependent instructions each; batches of regular tiles in one binade now compose their maps by a warp scan, above:
        // -8 %) and these additions.
        double cc = __shfl_sync(0xffffffffu, c, 0);
        double vnext = lo + lane < hi ? src(seg, lo + lane) : 0.0;
        for (int i0 = lo; i0 < hi; i0 += 32) {
          const int i = i0 + lane, m = min(32, hi - i0);
          const double v = vnext;
          vnext = i + 32 < hi ? src(seg, i + 32) : 0.0;     // the next round's value (load, exp, divisions) under this round's additions
          double mine = 0.0;
#pragma unroll
          for (int k = 0; k < 32; k++) {
            const double vk = __shfl_sync(0xffffffffu, v, k);
            if (k < m) {
              cc = (i0 + k == 0) ? vk : cc + vk;  // c[0] = p[0] (src/resampling.cpp:25)
              if (k == lane) mine = cc;
            }
          }
          if (cdf && i < hi) cdf[(size_t)seg * cdf_stride + i] = mine;
        }
        c = cc;
      }
    }
    __syncwarp();
    if (lane < cnt) {
      cstart[(size_t)seg * ntiles + t0 + lane] = s_cstart[lane];
      used_be[(size_t)seg * ntiles + t0 + lane] = s_use[lane];
    }
    __syncwarp();
  }
  if (lane == 0) {
    total_out[seg] = c;
    if (n_serial) n_serial[seg] = serial;
  }
  (void)s_c;
}

// ---- K4d: exact cdf inside the regular tiles -------------------------------------------------
template <typename Src>
__global__ void __launch_bounds__(RS_THREADS) k_tile_exact(Src src, int n, const int* __restrict__ n_per_seg,
                                                          int ntiles, const double* __restrict__ cstart,
                                                          const int* __restrict__ used_be,
                                                          double* __restrict__ cdf, size_t cdf_stride,
                                                          const int* __restrict__ enable) {
  __shared__ i64 fsm[2][RS_THREADS / 32];
  int seg = blockIdx.x, tile = blockIdx.y;   // segment in grid.x (no 65535 limit)
  if (!seg_on(enable, seg)) return;
  int nn = n_per_seg ? n_per_seg[seg] : n;
  if (tile * RS_TILE >= nn) return;
  int be = used_be[(size_t)seg * ntiles + tile];
  if (be < 0) return;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  i64 C0 = to_units(cstart[(size_t)seg * ntiles + tile], be);
  int base = tile * RS_TILE + threadIdx.x * RS_IPT;
  ParFn loc[RS_IPT];
  ParFn f = parfn_identity();
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    if (base + k < nn) f = parfn_compose(f, parfn_element(src(seg, base + k), be));
    loc[k] = f;
  }
  ParFn inc = f;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    ParFn g; g.a0 = shfl_up_i64(inc.a0, o); g.a1 = shfl_up_i64(inc.a1, o);
    if (lane >= o) inc = parfn_compose(g, inc);
  }
  if (lane == 31) { fsm[0][wid] = inc.a0; fsm[1][wid] = inc.a1; }
  // exclusive map of this thread inside the warp
  ParFn ex; ex.a0 = shfl_up_i64(inc.a0, 1); ex.a1 = shfl_up_i64(inc.a1, 1);
  if (lane == 0) ex = parfn_identity();
  __syncthreads();
  ParFn pre = parfn_identity();
  for (int w = 0; w < wid; w++) { ParFn x; x.a0 = fsm[0][w]; x.a1 = fsm[1][w]; pre = parfn_compose(pre, x); }
  pre = parfn_compose(pre, ex);
  i64 Cthread = parfn_apply(pre, C0);
#pragma unroll
  for (int k = 0; k < RS_IPT; k++) {
    int i = base + k;
    if (i < nn) cdf[(size_t)seg * cdf_stride + i] = from_units(parfn_apply(loc[k], Cthread), be);
  }
}

// ---- K5 / K6: index search --------------------------------------------------------------------
// first j in [0, n-1] with cdf[j] >= pos, clamped to n-1
__device__ __forceinline__ int lower_bound_clamped(const double* __restrict__ cdf, int n, double pos) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = lo + ((hi - lo) >> 1);
    if (cdf[mid] < pos) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// uniform sources: injected buffer (u[seg*u_stride + i], systematic: i = 0) or Philox
struct USrcBuf {
  const double* u; size_t stride;
  __device__ __forceinline__ double operator()(int seg, int i) const { return u[(size_t)seg * stride + i]; }
};

// position of output slot i (src/resampling.cpp:28,55); multinomial: the uniform itself
template <typename USrc>
__device__ __forceinline__ double resample_pos(const USrc& us, int fn, int seg, int i, int n) {
  if (fn == 1 /*systematic*/) return ((double)i + us(seg, 0)) / (double)n;
  double u = us(seg, i);
  if (fn == 0 /*stratified*/) return ((double)i + u) / (double)n;
  return u;
}

// standalone search: writes 1-based ancestors
template <typename USrc>
__global__ void __launch_bounds__(256) k_search(USrc us, int fn, int n, const int* __restrict__ n_per_seg,
                                               const double* __restrict__ cdf, size_t cdf_stride,
                                               int* __restrict__ idx1, size_t idx_stride,
                                               const int* __restrict__ enable) {
  int seg = blockIdx.x;
  if (!seg_on(enable, seg)) return;
  int nn = n_per_seg ? n_per_seg[seg] : n;
  const double* c = cdf + (size_t)seg * cdf_stride;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < nn; i += gridDim.y * blockDim.x) {
    double pos = resample_pos(us, fn, seg, i, nn);
    idx1[(size_t)seg * idx_stride + i] = lower_bound_clamped(c, nn, pos) + 1;
  }
}

// serial reference-order scan on the device (diagnostics / cross-check of the exact path)
template <typename Src>
__global__ void k_serial_scan(Src src, int n, double* __restrict__ cdf, double* __restrict__ total) {
  if (blockIdx.x || threadIdx.x) return;
  double c = 0.0;
  for (int i = 0; i < n; i++) { double v = src(0, i); c = (i == 0) ? v : c + v; if (cdf) cdf[i] = c; }
  if (total) *total = c;
}

}  // namespace bssm
