// simt_emu.h -- a small SIMT emulation for CPU logic tests of CUDA kernels that use block barriers, warp shuffles,
// shared memory and tickets (test infrastructure; not part of the product).
//
// A kernel's source text is compiled by g++ as ordinary C++: the CUDA qualifiers are defined away, `__shared__`
// becomes `static` (blocks run ONE AT A TIME, so a static local is exactly one block's shared memory), threadIdx /
// blockIdx are globals the scheduler sets, and every CUDA thread of the running block is a fiber (ucontext) on the one
// host thread, run round-robin until it reaches a barrier.  __syncthreads() is a generation barrier over the block, a
// warp shuffle is an exchange through a per-warp buffer between two warp barriers (all 32 lanes must take part, as
// with a full mask on the device); a round in which no barrier opens and no fiber finishes is a deadlock and aborts
// with a message.  Deterministic.  Blocks run in ascending order by default or in a caller-chosen order -- a kernel
// whose blocks meet only through tickets and fences must give the same answer for any order.
//
// What this checks: indexing, barrier placement, scan / expansion / merge logic and the arithmetic in IEEE double
// (compile with -ffp-contract=off: the device's FMA contraction is the one thing not reproduced).  What it does not:
// memory-model races between blocks, performance, the SFU approximations of the throughput precision (libm here).
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <functional>
#include <thread>
#include <ucontext.h>
#include <vector>

#define BSSM_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __grid_constant__
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct emu_dim3 { unsigned int x = 1, y = 1, z = 1; };
struct emu_idx3 { unsigned int x = 0, y = 0, z = 0; };
static thread_local emu_idx3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;

struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct alignas(8) uint2 { unsigned int x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }

// ---- the running block: fibers, generation barriers ----
struct EmuBarrier { int expected = 0, arrived = 0; unsigned int gen = 0; };
struct EmuFiber { ucontext_t ctx; bool done = false; };
struct EmuBlock {
  EmuBarrier bar;
  EmuBarrier nbar[16];            // named barriers (bar.sync / bar.arrive with an id and a thread count)
  std::vector<EmuBarrier> wbar;
  std::vector<uint64_t> xch;
  int orv = 0;
};
static thread_local EmuBlock* emu_block = nullptr;
static thread_local std::vector<EmuFiber> emu_fibers;
static thread_local ucontext_t emu_sched_ctx;
static thread_local int emu_cur = -1;
static thread_local long emu_progress = 0;
static thread_local std::vector<unsigned char> emu_dyn_smem;   // `extern __shared__` of the running block
static std::function<void()> emu_body;                        // one kernel at a time
static inline unsigned char* emu_dynamic_smem() { return emu_dyn_smem.data(); }

static inline void emu_yield() { swapcontext(&emu_fibers[emu_cur].ctx, &emu_sched_ctx); }
// a thread polling memory another block writes (cooperative launches): let the other fibers and blocks run
static inline void emu_poll_yield() { emu_yield(); }
static inline void emu_open(EmuBarrier& b) { b.arrived = 0; b.gen++; emu_progress++; }
static inline void emu_wait(EmuBarrier& b) {
  const unsigned int g = b.gen;
  if (++b.arrived >= b.expected) { emu_open(b); return; }
  while (b.gen == g) emu_yield();
}
static inline void emu_drop(EmuBarrier& b) {   // a fiber that has left the kernel no longer takes part
  b.expected--;
  if (b.expected > 0 && b.arrived >= b.expected) emu_open(b);
}
static void emu_trampoline() {
  emu_body();
  const int t = emu_cur;
  emu_fibers[t].done = true;
  emu_drop(emu_block->wbar[t >> 5]);
  emu_drop(emu_block->bar);
  emu_progress++;
  swapcontext(&emu_fibers[t].ctx, &emu_sched_ctx);
}

static inline void __syncthreads() { emu_wait(emu_block->bar); }
// PTX named barriers: `nthreads` threads take part, some by waiting (sync), some by only announcing themselves (arrive)
static inline void bar_sync_named(int id, int nthreads) { EmuBarrier& b = emu_block->nbar[id]; b.expected = nthreads; emu_wait(b); }
static inline void bar_arrive_named(int id, int nthreads) {
  EmuBarrier& b = emu_block->nbar[id];
  b.expected = nthreads;
  if (++b.arrived >= b.expected) emu_open(b);
}
static inline int __syncthreads_or(int pred) {
  if (pred) emu_block->orv = 1;
  emu_wait(emu_block->bar);
  const int r = emu_block->orv;
  emu_wait(emu_block->bar);
  if (threadIdx.x == 0) emu_block->orv = 0;
  emu_wait(emu_block->bar);
  return r;
}
template <typename T> static inline T emu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  EmuBlock& B = *emu_block;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  B.xch[(size_t)w * 32 + lane] = bits;
  emu_wait(B.wbar[w]);
  const uint64_t r = B.xch[(size_t)w * 32 + (src_lane & 31)];
  emu_wait(B.wbar[w]);
  T out;
  memcpy(&out, &r, sizeof(T));
  return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return emu_shfl(v, (int)(threadIdx.x & 31) ^ o); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int o) {
  const int lane = threadIdx.x & 31;
  return emu_shfl(v, lane >= o ? lane - o : lane);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int o) {
  const int lane = threadIdx.x & 31;
  return emu_shfl(v, lane + o < 32 ? lane + o : lane);
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_wait(emu_block->wbar[threadIdx.x >> 5]); }
static inline unsigned int __ballot_sync(unsigned, int pred) {
  EmuBlock& B = *emu_block;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  B.xch[(size_t)w * 32 + lane] = pred ? 1u : 0u;
  emu_wait(B.wbar[w]);
  unsigned int m = 0;
  for (int l = 0; l < B.wbar[w].expected; l++) if (B.xch[(size_t)w * 32 + l]) m |= 1u << l;
  emu_wait(B.wbar[w]);
  return m;
}
static inline int __all_sync(unsigned m, int pred) {
  const unsigned int b = __ballot_sync(m, pred);
  const int n = emu_block->wbar[threadIdx.x >> 5].expected;
  return b == (n >= 32 ? 0xffffffffu : ((1u << n) - 1u));
}
static inline int __popc(unsigned int v) { return __builtin_popcount(v); }
static inline int __reduce_max_sync(unsigned, int v) {
  EmuBlock& B = *emu_block;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  B.xch[(size_t)w * 32 + lane] = (uint64_t)(int64_t)v;
  emu_wait(B.wbar[w]);
  int m = v;
  for (int l = 0; l < B.wbar[w].expected; l++) { const int o = (int)(int64_t)B.xch[(size_t)w * 32 + l]; m = o > m ? o : m; }
  emu_wait(B.wbar[w]);
  return m;
}

static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline long long clock64() { return 0; }

// ---- bit casts and the device math names ----
static inline float __uint_as_float(unsigned int v) { float f; memcpy(&f, &v, 4); return f; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline unsigned int __float_as_uint(float f) { unsigned int v; memcpy(&v, &f, 4); return v; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
// glibc's <math.h> already declares __sinf, __cosf, __sincosf, __logf, __expf (its internal aliases of the libm functions)
#define __sinf(x) sinf(x)
#define __cosf(x) cosf(x)
#define __sincosf(x, s, c) sincosf((x), (s), (c))
#define __logf(x) logf(x)
#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// ---- launch ----
// one block on the calling OS thread: its threads as round-robin fibers.  coop: other blocks run on other OS threads and
// this block may be waiting for them, so a round without progress is not a deadlock (only a long silence is)
static void emu_run_block(unsigned int b, unsigned int block, size_t smem_bytes, bool coop) {
  constexpr size_t STACK = 256 * 1024;
  static thread_local std::vector<char> stacks;
  if (stacks.size() < STACK * block) stacks.resize(STACK * block);
  if (emu_fibers.size() < block) emu_fibers.resize(block);
  emu_dyn_smem.assign(smem_bytes, 0xCD);          // exact size (AddressSanitizer sees overruns), garbage contents
  blockIdx.x = b;
  EmuBlock blk;
  const int nw = (int)(block + 31) / 32;
  blk.bar.expected = (int)block;
  blk.wbar.resize(nw);
  for (int w = 0; w < nw; w++) blk.wbar[w].expected = (int)block - 32 * w < 32 ? (int)block - 32 * w : 32;
  blk.xch.assign((size_t)nw * 32, 0);
  emu_block = &blk;
  for (unsigned int t = 0; t < block; t++) {
    EmuFiber& f = emu_fibers[t];
    f.done = false;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stacks.data() + STACK * t;
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = &emu_sched_ctx;
    makecontext(&f.ctx, emu_trampoline, 0);
  }
  unsigned int remaining = block;
  long idle_rounds = 0;
  while (remaining) {
    const long before = emu_progress;
    for (unsigned int t = 0; t < block; t++) {
      if (emu_fibers[t].done) continue;
      emu_cur = (int)t;
      threadIdx.x = t;
      swapcontext(&emu_sched_ctx, &emu_fibers[t].ctx);
      if (emu_fibers[t].done) remaining--;
    }
    if (remaining && emu_progress == before) {
      if (coop && ++idle_rounds < 20000000L) { std::this_thread::yield(); continue; }
      fprintf(stderr, "simt_emu: deadlock in block %u (%u threads waiting at barriers that cannot open)\n", b, remaining);
      abort();
    }
    idle_rounds = 0;
  }
  emu_block = nullptr;
}
// ordinary launch: blocks one at a time, in `order` or ascending
template <typename Body> static void emu_launch(unsigned int grid, unsigned int block, Body body, const std::vector<unsigned int>* order = nullptr,
                                                size_t smem_bytes = 0) {
  gridDim.x = grid;
  blockDim.x = block;
  emu_body = body;
  for (unsigned int bi = 0; bi < grid; bi++) emu_run_block(order ? (*order)[bi] : bi, block, smem_bytes, false);
}
// ordinary launch of a two-dimensional grid (blockIdx.x fastest)
template <typename Body> static void emu_launch2d(unsigned int gx, unsigned int gy, unsigned int block, Body body) {
  gridDim.x = gx; gridDim.y = gy;
  blockDim.x = block;
  emu_body = body;
  for (unsigned int by = 0; by < gy; by++)
    for (unsigned int bx = 0; bx < gx; bx++) { blockIdx.y = by; emu_run_block(bx, block, 0, false); }
  gridDim.y = 1; blockIdx.y = 0;
}
// cooperative launch: all blocks co-resident, one OS thread per block
template <typename Body> static void emu_launch_cooperative(unsigned int grid, unsigned int block, size_t smem_bytes, Body body) {
  gridDim.x = grid;
  blockDim.x = block;
  emu_body = body;
  std::vector<std::thread> blocks;
  for (unsigned int b = 0; b < grid; b++) blocks.emplace_back([=] { emu_run_block(b, block, smem_bytes, true); });
  for (auto& th : blocks) th.join();
}
