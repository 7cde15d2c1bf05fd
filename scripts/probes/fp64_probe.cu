// fp64 latency / throughput probe (one SM): how expensive are DADD / DFMA / DSETP chains on this part?
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void chain(T* out, long long* cyc, int iters, int ilp) {
  T a0 = (T)threadIdx.x * (T)1e-3, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
  const T b = (T)1.0000001, c = (T)1e-9;
  __syncthreads();
  long long t0 = clock64();
  if (ilp == 1) { for (int i = 0; i < iters; i++) a0 = a0 * b + c; }
  else { for (int i = 0; i < iters; i++) { a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c; } }
  long long t1 = clock64();
  __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <typename T> void run(const char* name) {
  T* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 4096;
  for (int ilp : {1, 4}) for (int warps : {1, 4, 8, 16, 32}) {
    chain<T><<<1, warps * 32>>>(out, cyc, iters, ilp); cudaDeviceSynchronize();
    chain<T><<<1, warps * 32>>>(out, cyc, iters, ilp); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / iters;
    printf("%s ilp=%d warps=%2d: %.1f cycles per loop iteration; %.2f warp-FMA/cycle/SM\n", name, ilp, warps, per, warps * ilp / per);
  }
}
int main() { run<float>("f32"); run<double>("f64"); return 0; }
