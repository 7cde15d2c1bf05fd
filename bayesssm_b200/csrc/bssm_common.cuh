// bssm_common.cuh -- shared device utilities: Philox4x32-10, noise keying, math wrappers.
// Self-contained (no host headers) so the same text compiles under nvcc and NVRTC.
#pragma once

#ifndef __CUDACC_RTC__
#include <stdint.h>
#else
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
#endif

#define BSSM_DEV __device__ __forceinline__
#define BSSM_HD __host__ __device__ __forceinline__

namespace bssm {

// ---- noise tags (the oracle restates the same table) ----
enum : uint32_t {
  TAG_INIT_Z = 1, TAG_TRANS_Z = 2, TAG_TRANS2_Z = 3, TAG_RESAMP_U = 4, TAG_RESAMP_AUX_U = 5,
  TAG_MOVE_Z = 6, TAG_MOVE_U = 7, TAG_TRANS_U = 8, TAG_TRANS2_U = 9, TAG_INIT_U = 10,
  TAG_TRANS_DYN = 11, TAG_TRANS2_DYN = 12,   // uniforms on demand (DynU): a transition that draws a data-dependent number of them
  TAG_THETA_Z = 16, TAG_THETA_U = 17
};
constexpr uint32_t T_INIT = 0xFFFFFFFFu;

struct uint4x { uint32_t w[4]; };

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011).
BSSM_HD uint4x philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  uint4x o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
  return o;
}

struct NoiseKey {
  uint32_t k0, k1;   // (seed_lo, seed_hi ^ run_id)
  uint32_t stream;   // filter / chain id
};
BSSM_HD NoiseKey make_key(uint64_t seed, uint32_t run_id, uint32_t stream) {
  NoiseKey k; k.k0 = (uint32_t)seed; k.k1 = (uint32_t)(seed >> 32) ^ run_id; k.stream = stream; return k;
}
// words for the quad containing particle `index` (index>>2)
BSSM_HD uint4x noise_quad(const NoiseKey& k, uint32_t t, uint32_t tag, uint32_t slot, uint32_t quad) {
  return philox4x32_10(quad, t, k.stream, tag | (slot << 8), k.k0, k.k1);
}
BSSM_HD double word_to_unit_f64(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }
// Uniforms on demand for ONE particle's transition (models with DYN_U, e.g. an exact Gillespie step): uniform k is word k & 3 of
// Philox(counter = (particle, t, stream, tag | (k >> 2) << 8)) -- the particle index itself in the first counter word, a tag of its own.
// Philox mode only (injected noise buffers have a fixed number of slots).
struct DynU {
  NoiseKey key; uint32_t t, tag, particle; uint4x q; int have;
  BSSM_HD DynU(const NoiseKey& k, uint32_t t_, uint32_t tag_, uint32_t particle_) : key(k), t(t_), tag(tag_), particle(particle_), have(-1) { q.w[0] = q.w[1] = q.w[2] = q.w[3] = 0u; }
  BSSM_HD double operator()(int k) {
    const int call = k >> 2;
    if (call != have) { q = philox4x32_10(particle, t, key.stream, tag | ((uint32_t)call << 8), key.k0, key.k1); have = call; }
    return word_to_unit_f64(q.w[k & 3]);
  }
};
// does a model draw its transition uniforms on demand?  (built-in and NVRTC user models alike: the member is optional)
template <typename M, typename = void> struct ModelDynU { static constexpr bool value = false; };
template <typename M> struct ModelDynU<M, decltype((void)M::DYN_U)> { static constexpr bool value = M::DYN_U; };
BSSM_HD float word_to_unit_f32(uint32_t w) { return ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// ---- packed fp32 pairs (sm_100a: add / mul / fma .f32x2 issue as ONE instruction for two lanes -- FADD2 / FMUL2 / FFMA2) ----
// Used where the streaming engine's tile loops do the same fp32 operation on neighbouring particles and are bound by issue
// slots.  Each lane is the IEEE operation of the scalar instruction.
#ifndef BSSM_EMU
struct F2 { unsigned long long v; };
BSSM_DEV F2 f2_make(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
BSSM_DEV void f2_get(F2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
BSSM_DEV F2 f2_add(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
BSSM_DEV F2 f2_mul(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
BSSM_DEV F2 f2_fma(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else   // CPU logic test (tests/simt_emu.h)
struct F2 { float lo, hi; };
inline F2 f2_make(float lo, float hi) { F2 r; r.lo = lo; r.hi = hi; return r; }
inline void f2_get(F2 a, float& lo, float& hi) { lo = a.lo; hi = a.hi; }
inline F2 f2_add(F2 a, F2 b) { return f2_make(a.lo + b.lo, a.hi + b.hi); }
inline F2 f2_mul(F2 a, F2 b) { return f2_make(a.lo * b.lo, a.hi * b.hi); }
inline F2 f2_fma(F2 a, F2 b, F2 c) { return f2_make(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
#endif

// ---- math wrappers templated on the state precision ----
template <typename Real> struct Math;
template <> struct Math<double> {
  static BSSM_DEV double unit(uint32_t w) { return word_to_unit_f64(w); }
  static BSSM_DEV void box_muller(uint32_t a, uint32_t b, double& n0, double& n1) {
    double u1 = word_to_unit_f64(a), u2 = word_to_unit_f64(b);
    double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincos(2.0 * 3.14159265358979323846 * u2, &s, &c);
    n0 = r * c; n1 = r * s;
  }
  static BSSM_DEV void box_muller4(const uint32_t* w, double* z) { box_muller(w[0], w[1], z[0], z[1]); box_muller(w[2], w[3], z[2], z[3]); }
  static BSSM_DEV double exp_(double x) { return exp(x); }
  static BSSM_DEV double log_(double x) { return log(x); }
  static BSSM_DEV double sin_(double x) { return sin(x); }
  static BSSM_DEV double cos_(double x) { return cos(x); }
  static BSSM_DEV double div_(double a, double b) { return a / b; }
  static BSSM_DEV double max_(double a, double b) { return b > a ? b : a; }
  static BSSM_DEV double ninf() { return -__longlong_as_double(0x7FF0000000000000LL); }
};
template <> struct Math<float> {
  // Throughput precision.  Uniforms come from the top 23 bits of a Philox word by bit assembly
  // (no int->float conversion on the XU pipe), logarithm / square root / sine / cosine / exp on the SFU.
  // (k + 0.5) * 2^-23 for the 23-bit k = w >> 9: one exact subtraction (1 + k 2^-23) - (1 - 2^-24)
  static BSSM_DEV float unit(uint32_t w) { return __uint_as_float(0x3F800000u | (w >> 9)) - 0.99999994f; }  // (0, 1)
  // SFU instructions in their flush-to-zero form: the default forms wrap every MUFU in a denormal test and two predicated
  // multiplies (issue slots, taken or not); no quantity here is a denormal that matters (weights below 2^-126 of the maximum)
#ifndef BSSM_EMU
  static BSSM_DEV float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
  static BSSM_DEV float ex2_(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
  static BSSM_DEV float lg2_(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
  static BSSM_DEV float rcp_(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else   // CPU logic test (tests/simt_emu.h)
  static BSSM_DEV float sqrt_approx(float x) { return sqrtf(x); }
  static BSSM_DEV float ex2_(float x) { return exp2f(x); }
  static BSSM_DEV float lg2_(float x) { return log2f(x); }
  static BSSM_DEV float rcp_(float x) { return 1.0f / x; }
#endif
  static BSSM_DEV void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    float u1 = unit(a), u2 = unit(b);
    float r = sqrt_approx(-1.3862943611198906f * lg2_(u1));   // sqrt(-2 ln u1)
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);  // argument in (0, 2pi): the SFU path is accurate to ~1e-6 here
    n0 = r * c; n1 = r * s;
  }
  // the four normals of a Philox quad (words 0, 1 -> pair 0; words 2, 3 -> pair 1), the two Box-Muller transforms side by side:
  // the same values as two box_muller() calls, the fp32 arithmetic between the SFU calls as packed pairs
  static BSSM_DEV void box_muller4(const uint32_t* w, float* z) {
    const F2 c = f2_make(-0.99999994f, -0.99999994f);
    float u1a, u1b, a0, a1;
    f2_get(f2_add(f2_make(__uint_as_float(0x3F800000u | (w[0] >> 9)), __uint_as_float(0x3F800000u | (w[2] >> 9))), c), u1a, u1b);
    f2_get(f2_mul(f2_add(f2_make(__uint_as_float(0x3F800000u | (w[1] >> 9)), __uint_as_float(0x3F800000u | (w[3] >> 9))), c),
                  f2_make(6.283185307179586f, 6.283185307179586f)), a0, a1);
    float l0, l1;
    f2_get(f2_mul(f2_make(lg2_(u1a), lg2_(u1b)), f2_make(-1.3862943611198906f, -1.3862943611198906f)), l0, l1);
    const float r0 = sqrt_approx(l0), r1 = sqrt_approx(l1);
    float s0, c0, s1, c1;
    __sincosf(a0, &s0, &c0); __sincosf(a1, &s1, &c1);
    f2_get(f2_mul(f2_make(r0, r0), f2_make(c0, s0)), z[0], z[1]);
    f2_get(f2_mul(f2_make(r1, r1), f2_make(c1, s1)), z[2], z[3]);
  }
  static BSSM_DEV float exp_(float x) { return ex2_(x * 1.4426950408889634f); }
  static BSSM_DEV float log_(float x) { return logf(x); }
  // reduce to [-pi, pi] with a two-term 2*pi (round-to-nearest by the 1.5*2^23 trick), then the SFU (abs. error < 1e-6)
  static BSSM_DEV float reduce_2pi(float x) {
    float k = (x * 0.15915494309189535f + 12582912.0f) - 12582912.0f;
    float r = fmaf(k, -6.2831854820251465f, x);
    return fmaf(k, 1.7484556e-7f, r);
  }
  static BSSM_DEV float sin_(float x) { return __sinf(reduce_2pi(x)); }
  static BSSM_DEV float cos_(float x) { return __cosf(reduce_2pi(x)); }
  // the same on a packed pair: the reduction's four fp32 operations once for two particles, the SFU call per lane
  static BSSM_DEV F2 reduce_2pi2(F2 x) {
    const F2 k = f2_add(f2_fma(x, f2_make(0.15915494309189535f, 0.15915494309189535f), f2_make(12582912.0f, 12582912.0f)), f2_make(-12582912.0f, -12582912.0f));
    return f2_fma(k, f2_make(1.7484556e-7f, 1.7484556e-7f), f2_fma(k, f2_make(-6.2831854820251465f, -6.2831854820251465f), x));
  }
  static BSSM_DEV F2 sin2_(F2 x) { float a, b; f2_get(reduce_2pi2(x), a, b); return f2_make(__sinf(a), __sinf(b)); }
  static BSSM_DEV F2 cos2_(F2 x) { float a, b; f2_get(reduce_2pi2(x), a, b); return f2_make(__cosf(a), __cosf(b)); }
  static BSSM_DEV float div_(float a, float b) { return a * rcp_(b); }
  static BSSM_DEV float max_(float a, float b) { return fmaxf(a, b); }   // a NaN operand never wins, like (b > a ? b : a) for a finite a
  static BSSM_DEV float ninf() { return -__int_as_float(0x7F800000); }
};

// R densities
template <typename Real> BSSM_DEV Real dnorm_log(Real x, Real mu, Real sigma, Real log_sigma) {
  Real z = Math<Real>::div_(x - mu, sigma);
  return -((Real)0.918938533204672741780329736406 + (Real)0.5 * z * z + log_sigma);
}
// throughput precision: the same value as two operations on the reciprocal (hoisted out of particle loops by the compiler)
template <> BSSM_DEV float dnorm_log<float>(float x, float mu, float sigma, float log_sigma) {
  const float z = (x - mu) * Math<float>::rcp_(sigma);
  return fmaf(z, -0.5f * z, -(0.918938533204672741780329736406f + log_sigma));
}
// dnorm_log<float> on a packed pair (the same operations per lane)
BSSM_DEV F2 dnorm_log2(float x, F2 mu, float sigma, float log_sigma) {
  const F2 z = f2_mul(f2_fma(mu, f2_make(-1.0f, -1.0f), f2_make(x, x)), f2_make(Math<float>::rcp_(sigma), Math<float>::rcp_(sigma)));
  const float c = -(0.918938533204672741780329736406f + log_sigma);
  return f2_fma(z, f2_mul(z, f2_make(-0.5f, -0.5f)), f2_make(c, c));
}
// does a model offer its transition / log-likelihood on packed pairs (throughput precision of the streaming engine)?
template <typename M, typename = void> struct ModelPacked { static constexpr bool value = false; };
template <typename M> struct ModelPacked<M, decltype((void)M::PACKED)> { static constexpr bool value = M::PACKED; };
template <typename Real> BSSM_DEV Real dpois_log(Real y, Real lambda) {
  if (lambda == (Real)0) return (y == (Real)0) ? (Real)0 : Math<Real>::ninf();
  return y * Math<Real>::log_(lambda) - lambda - (Real)lgamma((double)y + 1.0);
}
// throughput precision: lgammaf instead of the fp64 lgamma (the same for every particle of an observation, but evaluated per
// thread: ~250 fp64 instructions against ~50)
template <> BSSM_DEV float dpois_log<float>(float y, float lambda) {
  if (lambda == 0.f) return (y == 0.f) ? 0.f : Math<float>::ninf();
  return y * Math<float>::log_(lambda) - lambda - lgammaf(y + 1.0f);
}
BSSM_DEV double binom_inversion(double nd, double p, double u);
// Binomial(n, 1 - exp(log_q)) by the same sequential inversion in fp32 (throughput precision of the integer-state models: the
// fp64 recurrence with its division per term is ~7x the instructions).  The chain-binomial rates arrive as log(1 - p) = -rate,
// so neither 1 - p nor log1p(-p) is formed in fp32.  Where (1 - p)^n would leave the fp32 range the fp64 routine serves.
BSSM_DEV float binom_inversion_f32(float nf, float log_q, float u) {
  const int n = (int)nf;
  if (n <= 0 || !(log_q < 0.f)) return 0.f;
  const float a = nf * log_q;                 // log of P(0)
  if (a < -80.f) return (float)binom_inversion((double)nf, -expm1((double)log_q), (double)u);
  const float q = Math<float>::exp_(log_q), r = (1.0f - q) * Math<float>::rcp_(q);
  float pmf = Math<float>::exp_(a), cdf = pmf;
  float kf = 0.f, left = nf;            // k and n - k as floats (exact): no conversions in the loop
  while (u > cdf && left > 0.f) {
    kf += 1.0f;
    pmf *= left * Math<float>::rcp_(kf) * r;
    left -= 1.0f;
    cdf += pmf;
  }
  return kf;
}
// Binomial(n, p) by sequential cdf inversion from one uniform; always double (matches the oracle)
BSSM_DEV double binom_inversion(double nd, double p, double u) {
  int n = (int)nd;
  if (n <= 0 || p <= 0) return 0.0;
  if (p >= 1) return (double)n;
  double q = 1.0 - p, r = p / q;
  double pmf = exp((double)n * log1p(-p));
  double cdf = pmf;
  int k = 0;
  while (u > cdf && k < n) {
    k++;
    pmf *= ((double)(n - k + 1) / (double)k) * r;
    cdf += pmf;
  }
  return (double)k;
}

}  // namespace bssm
