// bssm_slots.cuh -- pieces shared by the persistent kernel (bssm_fast.cuh) and the streaming engine
// (bssm_stream.cuh): fp64 warp reductions / scans and the closed-form output-slot counter of stratified and
// systematic resampling (src/resampling.cpp:16-66).  No host headers: the same text compiles under NVRTC.
#pragma once
#include "bssm_common.cuh"

namespace bssm {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}
__device__ __forceinline__ double warp_incl_scan_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
  return v;
}

// Output-slot bookkeeping of one resampling step.  Position of slot i: (i + U_i) / n (stratified,
// U_i = Philox word of slot i) or (i + U) / n (systematic).  count_le(c) = #{ i : pos_i <= c }
// = first slot whose position exceeds c.  With t = c*n and i = floor(t): slots below i have
// i' + U < i' + 1 <= t, slots above have i' >= i + 1 > t, so only slot i needs a look.
// rare path (slot outside the staged window): kept out of line so the hot loop stays small (+3 %, A/B)
static __device__ __noinline__ unsigned int philox_word_slow(NoiseKey key, unsigned int obs, int i) {
  uint4x q = noise_quad(key, obs, TAG_RESAMP_U, 0u, (unsigned int)i >> 2);
  return q.w[i & 3];
}
struct SlotCounter {
  NoiseKey key; unsigned int obs; int fn; int n; unsigned int w_sys;
  const unsigned int* s_u; int u_base, u_cap;   // staged Philox words for slots [u_base, u_base + u_cap)
  __device__ __forceinline__ unsigned int word_of(int i) const {
    if (fn == 1) return w_sys;
    unsigned int k = (unsigned int)(i - u_base);
    if (k < (unsigned int)u_cap) return s_u[k];
    return philox_word_slow(key, obs, i);
  }
  __device__ __forceinline__ int count_slots(double t) const {   // exact rule (fp64), t = c * n: a position in units of slots
    if (!(t > 0.0)) return 0;
    if (t >= (double)n) return n;
    int i = (int)t;
    return i + (((double)i + word_to_unit_f64(word_of(i))) <= t ? 1 : 0);
  }
  __device__ __forceinline__ int count_le(double c) const { return count_slots(c * (double)n); }
};

}  // namespace bssm
