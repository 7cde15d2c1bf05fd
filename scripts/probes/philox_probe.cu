// philox_probe.cu -- what does one Philox4x32-10 call cost on the B200?  14 warps per SM (the persistent kernel's worker
// geometry), 4 independent calls per thread and iteration (as gen_normals issues them), three formulations of the round
// multiply: 0 = 32x32->64 product (IMAD.WIDE), 1 = separate mul.hi / mul.lo, 2 = as 0 plus Box-Muller on the SFU.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o philox_probe philox_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
template <int MODE> __device__ __forceinline__ void philox(unsigned& c0, unsigned& c1, unsigned& c2, unsigned& c3, unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    unsigned hi0, lo0, hi1, lo1;
    if (MODE == 1) {
      asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi0) : "r"(0xD2511F53u), "r"(c0)); asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo0) : "r"(0xD2511F53u), "r"(c0));
      asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi1) : "r"(0xCD9E8D57u), "r"(c2)); asm("mul.lo.u32 %0, %1, %2;" : "=r"(lo1) : "r"(0xCD9E8D57u), "r"(c2));
    } else {
      hi0 = __umulhi(0xD2511F53u, c0); lo0 = 0xD2511F53u * c0; hi1 = __umulhi(0xCD9E8D57u, c2); lo1 = 0xCD9E8D57u * c2;
    }
    unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float unit(unsigned w) { return (__uint_as_float(0x3F800000u | (w >> 9)) - 1.0f) + 5.9604645e-8f; }
template <int MODE> __global__ void __launch_bounds__(448, 1) k(int iters, unsigned* out, long long* cyc) {
  unsigned acc = 0;
  float facc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int h = 0; h < 4; h++) {
      unsigned c0 = threadIdx.x * 4 + h, c1 = it, c2 = blockIdx.x, c3 = 2;
      philox<MODE>(c0, c1, c2, c3, 1405u, 7u);
      if (MODE == 2) {
        float r0 = sqrtf(-2.0f * __logf(unit(c0))), r1 = sqrtf(-2.0f * __logf(unit(c2)));
        float s0, q0, s1, q1;
        __sincosf(6.2831853f * unit(c1), &s0, &q0); __sincosf(6.2831853f * unit(c3), &s1, &q1);
        facc += r0 * s0 + r0 * q0 + r1 * s1 + r1 * q1;
      } else acc ^= c0 ^ c1 ^ c2 ^ c3;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(facc);
}
int main() {
  const int G = 148, iters = 2000;
  unsigned* out; long long* cyc;
  cudaMalloc(&out, G * 448 * 4); cudaMalloc(&cyc, G * 8);
  for (int mode = 0; mode < 3; mode++) {
    if (mode == 0) k<0><<<G, 448>>>(iters, out, cyc); else if (mode == 1) k<1><<<G, 448>>>(iters, out, cyc); else k<2><<<G, 448>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    std::vector<long long> h(G);
    cudaMemcpy(h.data(), cyc, G * 8, cudaMemcpyDeviceToHost);
    double a = 0; for (auto v : h) a += (double)v; a /= G;
    printf("mode %d: %.0f cycles per iteration of 4 calls per thread (14 warps/SM) = %.1f cycles per call per warp-scheduler slot; %s\n", mode, a / iters, a / iters / 4 / 3.5, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
