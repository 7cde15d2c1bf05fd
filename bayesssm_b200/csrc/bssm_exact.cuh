// bssm_exact.cuh -- the arithmetic that lets a PARALLEL scan reproduce, bit for bit, the
// SEQUENTIAL double-precision running sum of the reference resamplers
//   total = 0; for i: total += w[i]                 (Rcpp::sum,    src/resampling.cpp:20,47)
//   c[0] = p[0]; c[i] = c[i-1] + p[i]               (Rcpp::cumsum, src/resampling.cpp:25,52)
//
// Idea.  While the running sum c stays inside one binade [2^e, 2^(e+1)) it is an integer
// multiple C of ulp = 2^(e-52), and  fl(c + p) = (C + Q + r) * ulp  with Q = floor(p/ulp) and
// r the round-to-nearest-even decision on the discarded fraction f = p/ulp - Q:
//   f < 1/2 -> r = 0;   f > 1/2 -> r = 1;   f == 1/2 -> r = (C + Q) & 1.
// So one addition is the map  C -> C + a[C & 1]  for a pair of integers (a0, a1); such maps
// compose into maps of the same form, and composition is associative, so they can be
// scanned in parallel exactly.  The few elements at which the running sum changes binade
// ("crossings", ~log2(n) of them) are applied with a real floating-point add in a short
// serial chain.  Which binade each element sees is taken from an ordinary approximate
// parallel scan and then VERIFIED against the exact values (2^52 <= C_start < 2^53 and
// C_end <= 2^53 per tile); on the (rare) failure the caller falls back to a serial kernel,
// so the result is always exactly the sequential one.
//
// Host+device, integer-only, no dependencies: the same text is compiled by nvcc for the
// kernels and by g++ for tests/host_exact_scan.cpp (logic test without a GPU).
#pragma once
#ifndef BSSM_HD
#ifdef __CUDACC__
#define BSSM_HD __host__ __device__ __forceinline__
#else
#define BSSM_HD inline
#endif
#endif

namespace bssm {

typedef long long i64;
typedef unsigned long long u64;

BSSM_HD u64 dbits(double x) {
#ifdef __CUDA_ARCH__
  return (u64)__double_as_longlong(x);
#else
  union { double d; u64 u; } v; v.d = x; return v.u;
#endif
}
BSSM_HD double bits_d(u64 b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((i64)b);
#else
  union { double d; u64 u; } v; v.u = b; return v.d;
#endif
}
// biased exponent (0 = zero/subnormal, 2047 = inf/nan) of a non-negative double
BSSM_HD int biased_exp(double x) { return (int)((dbits(x) >> 52) & 0x7FF); }

// C -> C + (C & 1 ? a1 : a0)
struct ParFn { i64 a0, a1; };
BSSM_HD ParFn parfn_identity() { ParFn f; f.a0 = 0; f.a1 = 0; return f; }
BSSM_HD i64 parfn_apply(const ParFn& f, i64 C) { return C + ((C & 1) ? f.a1 : f.a0); }
// (g then f)
BSSM_HD ParFn parfn_compose(const ParFn& g, const ParFn& f) {
  ParFn h;
  h.a0 = g.a0 + ((g.a0 & 1) ? f.a1 : f.a0);
  h.a1 = g.a1 + (((g.a1 + 1) & 1) ? f.a1 : f.a0);
  return h;
}

// Map of "add p" for a running sum in the binade with biased exponent be (1 <= be <= 2046).
// p >= 0 finite.  If p is too large for the sum to stay in the binade the returned
// increment is huge, which makes the caller's C_end <= 2^53 verification fail.
BSSM_HD ParFn parfn_element(double p, int be) {
  u64 b = dbits(p);
  int bp = (int)((b >> 52) & 0x7FF);
  u64 frac = b & 0xFFFFFFFFFFFFFull;
  u64 mp = bp ? (frac | (1ull << 52)) : frac;  // p = mp * 2^(max(bp,1) - 1075)
  int s = be - (bp ? bp : 1);                  // p / ulp = mp * 2^-s
  ParFn f;
  if (s <= 0) {
    int up = -s; if (up > 9) up = 9;
    f.a0 = f.a1 = (i64)(mp << up);
    return f;
  }
  if (s >= 54) { f.a0 = f.a1 = 0; return f; }
  u64 Q = mp >> s;
  u64 rem = mp & ((1ull << s) - 1ull);
  u64 half = 1ull << (s - 1);
  if (rem > half) { f.a0 = f.a1 = (i64)(Q + 1); }
  else if (rem < half) { f.a0 = f.a1 = (i64)Q; }
  else { f.a0 = (i64)(Q + (Q & 1)); f.a1 = (i64)(Q + ((Q & 1) ^ 1)); }
  return f;
}

// running sum <-> integer multiple of ulp(be).  to_units requires 2^(be-1023) <= c <= 2^(be-1022).
BSSM_HD i64 to_units(double c, int be) {
  u64 b = dbits(c);
  int bc = (int)((b >> 52) & 0x7FF);
  u64 frac = b & 0xFFFFFFFFFFFFFull;
  if (bc == be) return (i64)(frac | (1ull << 52));
  if (bc == be + 1 && frac == 0) return (i64)(1ull << 53);
  return -1;  // not in the binade: verification failure
}
BSSM_HD double from_units(i64 C, int be) {  // 2^52 <= C <= 2^53
  if (C == (i64)(1ull << 53)) return bits_d((u64)(be + 1) << 52);
  return bits_d(((u64)be << 52) | ((u64)C & 0xFFFFFFFFFFFFFull));
}
BSSM_HD bool units_ok_start(i64 C) { return C >= (i64)(1ull << 52) && C < (i64)(1ull << 53); }
BSSM_HD bool units_ok_end(i64 C) { return C >= (i64)(1ull << 52) && C <= (i64)(1ull << 53); }

}  // namespace bssm
