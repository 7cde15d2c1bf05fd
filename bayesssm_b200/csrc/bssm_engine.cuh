// bssm_engine.cuh -- host-side context, scratch memory and launch helpers shared by the
// translation units of libbayesssm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/bayesssm_b200.h"
#include "bssm_filter.cuh"
#include "bssm_resample.cuh"

namespace bssm {

void set_error(const char* fmt, ...);

#define BSSM_CK(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      bssm::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      return BSSM_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)
#define BSSM_TRY(call)            \
  do {                            \
    int s__ = (call);             \
    if (s__ != BSSM_OK) return s__; \
  } while (0)

// growable device scratch buffers owned by the context
enum ScratchSlot {
  SL_RS_PART = 0, SL_RS_REC, SL_RS_CSTART, SL_RS_USED, SL_RS_TOTAL, SL_RS_NSERIAL, SL_RS_STATUS, SL_RS_PREF,
  SL_API_W, SL_API_U, SL_API_IDX, SL_API_CDF,
  SL_F_XA, SL_F_XB, SL_F_LW, SL_F_LWAUX, SL_F_AUXG, SL_F_PART, SL_F_CDF, SL_F_SCAL_D, SL_F_SCAL_I,
  SL_F_ESS, SL_F_SEST, SL_F_LLH, SL_F_PH, SL_F_WH, SL_F_ANC, SL_F_ANCA, SL_F_THETA, SL_F_Y, SL_F_OBS,
  SL_F_IDS, SL_F_NOISE0, /* 10 noise slots */
  SL_F_NOISE_LAST = SL_F_NOISE0 + 9,
  SL_P_BASE, /* PMMH slots */
  SL_P_LAST = SL_P_BASE + 23,
  SL_FAST_BASE,
  SL_FAST_LAST = SL_FAST_BASE + 7,
  SL_ST_BASE,   /* streaming engine */
  SL_ST_LAST = SL_ST_BASE + 13,
  SL_DG_BASE,   /* MCMC diagnostics (bssm_diag.cu) */
  SL_DG_LAST = SL_DG_BASE + 5,
  SL_COUNT
};

struct Scratch { void* p = nullptr; size_t cap = 0; };

// handles of the model-dependent kernels of the general engine (built-in: function addresses;
// NVRTC user model: cudaKernel_t from the compiled library)
struct ModelKernels { void *init = nullptr, *weight = nullptr, *post = nullptr; bool has_aux = false, has_move = false; };
// handles of the streaming engine's kernels (bssm_stream.cuh): k_st_init, k_st_step, k_st_resample, k_st_flush
struct StreamKernels { void *init = nullptr, *step = nullptr, *resample = nullptr, *flush = nullptr, *chain = nullptr; };   // chain: the chain-persistent kernel (built-in models)
constexpr int BSSM_USER_MODEL_BASE = 1000;
struct UserModelInfo {
  void* library = nullptr;   // cudaLibrary_t
  ModelKernels k32, k64;
  StreamKernels s32[2], s64[2];   // streaming engine, [0] 256 / [1] 128 threads per block (valid when stream_ok)
  bool stream_ok = false;    // 1-D state, one normal per init / transition, no uniforms
  int dims[13];              // D, NTHETA, NCONST, NZ_INIT, NU_INIT, NZ_TRANS, NU_TRANS, NZ_MOVE, NU_MOVE, HAS_AUX, HAS_MOVE, NPAR, DYN_U
};

}  // namespace bssm

struct bssm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t copy_stream = nullptr;            // history rows travel on it (bssm_engine.cu: hist_row_out)
  cudaEvent_t ev_row[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  cudaStream_t aux_stream = nullptr;             // streaming engine, multinomial: the positions of the next observation are laid out here
  cudaEvent_t ev_mn_ready[2] = {nullptr, nullptr}, ev_mn_free[2] = {nullptr, nullptr}, ev_mn_start = nullptr;
  cudaDeviceProp prop;
  int64_t launches = 0;
  bssm::Scratch scratch[bssm::SL_COUNT];
  std::string compile_log;
  std::vector<bssm::UserModelInfo> user_models;
  void* nccl_comm = nullptr;   // ncclComm_t of the shard group (bssm_shard.cu)
  int shard_rank = 0, shard_world = 1;
  // peer-memory exchange of the sharded filter (bssm_shard_peer_export / _attach): every rank's inbox, mapped here with CUDA IPC
  void* peer_inbox = nullptr;              // this rank's own allocation
  void* peer_ptr[16] = {nullptr};          // [world] rank g's inbox as this process sees it (peer_ptr[shard_rank] == peer_inbox)
  int peer_on = 0;
  unsigned long long peer_seq = 1;         // next unused sequence number (advances identically on every rank: the calls are collective)
};

namespace bssm {

int scratch_get(bssm_ctx* ctx, int slot, size_t bytes, void** out);
template <typename T> inline int scratch(bssm_ctx* ctx, int slot, size_t count, T** out) {
  void* p = nullptr;
  int st = scratch_get(ctx, slot, count * sizeof(T), &p);
  *out = (T*)p;
  return st;
}
int check_launch(bssm_ctx* ctx, const char* what);

#define BSSM_LAUNCH(ctx, what)                       \
  do {                                               \
    (ctx)->launches++;                               \
    BSSM_TRY(bssm::check_launch((ctx), what));       \
  } while (0)

// cdf of `nseg` weight vectors (see bssm_resample.cuh).  Src yields the weights, SrcN the
// weights divided by the exact total (only used when exact != 0).
struct RsArgs {
  int nseg, n;             // n = stride / max length
  const int* n_per;        // [nseg] or nullptr
  const int* enable;       // [nseg] or nullptr
  double* cdf; size_t cdf_stride;
  int* status;             // [nseg] validation result (atomicMax) or nullptr
  int validate;
  int exact;
  long long* n_serial;     // [nseg] or nullptr: elements the chain walked serially (diagnostic)
  double* total;           // [nseg] out (exact: sequential sum of the raw weights)
};

// ---- batched filter runs (bssm_engine.cu), reused by the PMMH driver -----------------------------
struct FilterLaunch {
  int model, precision, resample_fn, exact, hist, T, engine;
  // return_particles: the caller's host buffers [C][T+1][d][N] / [C][T+1][N]; rows stream there while the filter runs
  double *h_particles_history = nullptr, *h_weights_history = nullptr;
};
int model_dims(bssm_ctx* ctx, int model, int* d, int* ntheta, int* nconst);
int resolve_engine(bssm_ctx* ctx, const FilterDev& f, const FilterLaunch& L, bool injected, bool want_anc);
int filter_setup(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, bool need_aux, bool want_anc, double** cdf_out, bool injected = false);
int filter_reset(bssm_ctx* ctx, FilterDev& f, const int* d_active);
int filter_enqueue(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, double* cdf);
// persistent bootstrap-filter kernel (bssm_fast.cu)
bool fast_supported(const FilterDev& f, const FilterLaunch& L);
int fast_filter_enqueue(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L);
// streaming engine (bssm_stream.cu) and its particle-sharded form (bssm_shard.cu)
struct ShardRun {
  int rank, world;
  int n_glob;        // particles of the whole filter
  long long goff0;   // this rank's initial slice [goff0, goff0 + nloc0)
  int nloc0;
  int cap;           // storage capacity of this rank (particles)
};
bool stream_supported(bssm_ctx* ctx, const FilterDev& f, const FilterLaunch& L);
int stream_filter_enqueue(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, const ShardRun* sh);
int shard_allgather(bssm_ctx* ctx, const ShardRun* sh, const void* d_send, void* d_recv, size_t bytes_per_rank);
// NVRTC user models (bssm_nvrtc.cu)
const UserModelInfo* user_model(bssm_ctx* ctx, int model_id);

}  // namespace bssm
