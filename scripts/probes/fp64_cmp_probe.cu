// is an fp64 compare-and-select chain (DSETP + SEL, as in a running max) slow on this part?  one SM, clock64
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_add(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-3, t = 1.0 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { t = t + 1e-9; a = a + t; }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_max(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-3, t = 1.0 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { t = t + 1e-9; a = t > a ? t : a; }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_imax(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-3, t = 1.0 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    t = t + 1e-9;
    long long ka = __double_as_longlong(a), kt = __double_as_longlong(t);
    ka ^= (ka >> 63) & 0x7FFFFFFFFFFFFFFFLL; kt ^= (kt >> 63) & 0x7FFFFFFFFFFFFFFFLL;
    a = kt > ka ? t : a;
  }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shflmax(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, a, o); a = t > a ? t : a; }
    a += 1e-9;
  }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shflsum(double* out, long long* cyc, int iters) {
  double a = threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    a *= 1e-3;
  }
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 64);
  const int iters = 2048;
  for (int warps : {1, 8, 14}) {
    long long h[5];
    k_add<<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); k_add<<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h[0], cyc, 8, cudaMemcpyDeviceToHost);
    k_max<<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); k_max<<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h[1], cyc, 8, cudaMemcpyDeviceToHost);
    k_imax<<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); k_imax<<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h[2], cyc, 8, cudaMemcpyDeviceToHost);
    k_shflmax<<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); k_shflmax<<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h[3], cyc, 8, cudaMemcpyDeviceToHost);
    k_shflsum<<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); k_shflsum<<<1, warps * 32>>>(out, cyc, iters); cudaMemcpy(&h[4], cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps=%2d cycles/iter: 2xDADD %.1f | DADD+fp64 max %.1f | DADD+int-key max %.1f | 5-step shuffle fp64 max %.1f | 5-step shuffle fp64 sum %.1f\n", warps,
           (double)h[0] / iters, (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters);
  }
  return 0;
}
