// bssm_shard.cu -- particle-sharded single filter (third row): one bootstrap filter
// whose particles are block-partitioned over the GPUs of one box, one process per GPU.
//
// Rank g holds the particles with global index in [goff_g, goff_g + nloc_g).  Per observation the
// ranks exchange ONE 64-byte record per filter (local max, sum e, sum e^2, sum e*x, pending state
// sum) with ncclAllGather over NVLink; every rank then derives -- from the same records, with the
// same expressions, in rank order -- the global max / normaliser / ESS / log-likelihood / resampling
// decision and the cdf offset of its own particles (the exclusive prefix of the per-rank weight
// totals).  Resampling is input-centric: rank g produces exactly the offspring of its own particles,
// i.e. the contiguous output slots [F(A_g / S), F(A_{g+1} / S)), which become its particles of the
// next step.  No particle ever crosses NVLink; the price is that nloc_g follows rank g's share of the
// weight, so each rank has storage for `capacity_factor` x N / G particles (BSSM_ERR_CAPACITY beyond).
// Philox streams are keyed by the GLOBAL particle index, so the result is the single-GPU result up to
// the floating-point summation order of the normaliser (tests/test_shard_gpu.py).
//
// The exchange has two forms.  Default after bssm_shard_init: ncclAllGather + a one-thread merge kernel between the two kernels
// of an observation.  After bssm_shard_peer_export / _attach (every rank maps every rank's 2 x world x 128-byte inbox with CUDA
// IPC): the merging block of k_st_step stores the record into every peer's inbox over NVLink, polls its own inbox and does the
// global bookkeeping itself -- compute and collective are one kernel, the host enqueues the same two launches per observation
// as on one GPU (programmatic dependent launch included).  Both forms sum the same records in the same order: identical results.
//
// NCCL is loaded at run time (dlopen), so the library has no link-time dependency on it; the unique
// id travels through whatever the host already has (torch.distributed in the Python mirror, a file
// or MPI under R).
#include "bssm_engine.cuh"
#include "bssm_stream.cuh"

#include <dlfcn.h>

namespace bssm {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_comm_init)(void**, int, nccl_uid, int);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_comm_destroy)(void*);
typedef const char* (*fn_errstr)(int);

static struct {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr; fn_comm_init comm_init = nullptr; fn_allgather allgather = nullptr;
  fn_comm_destroy comm_destroy = nullptr; fn_errstr errstr = nullptr;
} g_nccl;

static int nccl_load(const char* path) {
  if (g_nccl.lib) return BSSM_OK;
  const char* cands[] = {path, getenv("BSSM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* c : cands) {
    if (!c || !*c) continue;
    g_nccl.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) { set_error("NCCL library not found (pass its path or set BSSM_NCCL_LIB): %s", dlerror()); return BSSM_ERR_NCCL; }
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.comm_init = (fn_comm_init)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.allgather = (fn_allgather)dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.comm_init || !g_nccl.allgather || !g_nccl.comm_destroy) {
    set_error("NCCL library lacks a required symbol");
    dlclose(g_nccl.lib); g_nccl.lib = nullptr;
    return BSSM_ERR_NCCL;
  }
  return BSSM_OK;
}
#define BSSM_NCCL(call)                                                                                        \
  do {                                                                                                         \
    int r__ = (call);                                                                                          \
    if (r__ != 0) { set_error("NCCL error %d (%s) in %s", r__, g_nccl.errstr ? g_nccl.errstr(r__) : "?", #call); return BSSM_ERR_NCCL; } \
  } while (0)

int shard_allgather(bssm_ctx* ctx, const ShardRun* sh, const void* d_send, void* d_recv, size_t bytes_per_rank) {
  if (sh->world == 1) {
    BSSM_CK(cudaMemcpyAsync(d_recv, d_send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
    return BSSM_OK;
  }
  if (!ctx->nccl_comm) { set_error("shard group not initialised (bssm_shard_init)"); return BSSM_ERR_NCCL; }
  BSSM_NCCL(g_nccl.allgather(d_send, d_recv, bytes_per_rank, 0 /* ncclInt8 */, ctx->nccl_comm, ctx->stream));
  return BSSM_OK;
}

// initial block partition of n particles over `world` ranks: boundaries at multiples of 4 (Philox quads)
static void shard_partition(int n, int world, int rank, long long* goff, int* nloc) {
  const long long quads = ((long long)n + 3) / 4;
  const long long q0 = quads * rank / world, q1 = quads * (rank + 1) / world;
  long long a = q0 * 4, b = q1 * 4;
  if (b > n) b = n;
  if (a > n) a = n;
  *goff = a; *nloc = (int)(b - a);
}

}  // namespace bssm

using namespace bssm;

extern "C" {

int bssm_shard_unique_id(const char* nccl_lib_path, void* id_out_128) {
  if (!id_out_128) { set_error("bssm_shard_unique_id: null output"); return BSSM_ERR_BAD_ARG; }
  BSSM_TRY(nccl_load(nccl_lib_path));
  nccl_uid id;
  BSSM_NCCL(g_nccl.get_uid(&id));
  memcpy(id_out_128, &id, sizeof(id));
  return BSSM_OK;
}

int bssm_shard_init(bssm_ctx* ctx, const char* nccl_lib_path, int rank, int world, const void* id_128) {
  if (!ctx || world < 1 || rank < 0 || rank >= world) { set_error("bssm_shard_init: bad rank / world"); return BSSM_ERR_BAD_ARG; }
  BSSM_CK(cudaSetDevice(ctx->device));
  if (ctx->nccl_comm) { g_nccl.comm_destroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  ctx->shard_rank = rank; ctx->shard_world = world;
  if (world == 1) return BSSM_OK;
  if (!id_128) { set_error("bssm_shard_init: null unique id"); return BSSM_ERR_BAD_ARG; }
  BSSM_TRY(nccl_load(nccl_lib_path));
  nccl_uid id;
  memcpy(&id, id_128, sizeof(id));
  BSSM_NCCL(g_nccl.comm_init(&ctx->nccl_comm, world, id, rank));
  return BSSM_OK;
}

// ---- peer-memory exchange (CUDA IPC over NVLink): the per-observation all-gather fused into k_st_step's tail ----
static_assert(BSSM_PEER_MAX_WORLD == sizeof(((bssm_ctx*)nullptr)->peer_ptr) / sizeof(void*), "bssm_ctx::peer_ptr holds one pointer per rank of the largest group");
static void peer_release(bssm_ctx* ctx) {
  for (int g = 0; g < BSSM_PEER_MAX_WORLD; g++) {
    if (ctx->peer_ptr[g] && ctx->peer_ptr[g] != ctx->peer_inbox) cudaIpcCloseMemHandle(ctx->peer_ptr[g]);
    ctx->peer_ptr[g] = nullptr;
  }
  if (ctx->peer_inbox) { cudaFree(ctx->peer_inbox); ctx->peer_inbox = nullptr; }
  ctx->peer_on = 0; ctx->peer_seq = 1;
  cudaGetLastError();
}

int bssm_shard_peer_export(bssm_ctx* ctx, void* handle_out_64) {
  if (!ctx || !handle_out_64) { set_error("bssm_shard_peer_export: null argument"); return BSSM_ERR_BAD_ARG; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  if (ctx->shard_world < 2 || ctx->shard_world > BSSM_PEER_MAX_WORLD) { set_error("bssm_shard_peer_export: needs a shard group of 2 .. %d ranks (bssm_shard_init)", BSSM_PEER_MAX_WORLD); return BSSM_ERR_BAD_ARG; }
  BSSM_CK(cudaSetDevice(ctx->device));
  BSSM_CK(cudaStreamSynchronize(ctx->stream));
  peer_release(ctx);
  const size_t bytes = (size_t)2 * ctx->shard_world * sizeof(StPeerSlot);
  BSSM_CK(cudaMalloc(&ctx->peer_inbox, bytes));
  BSSM_CK(cudaMemset(ctx->peer_inbox, 0, bytes));
  BSSM_CK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, ctx->peer_inbox);
  if (e != cudaSuccess) {
    set_error("bssm_shard_peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    peer_release(ctx);
    return BSSM_ERR_CUDA;
  }
  memcpy(handle_out_64, &h, sizeof(h));
  return BSSM_OK;
}

int bssm_shard_peer_attach(bssm_ctx* ctx, const void* handles_world_x_64) {
  if (!ctx || !handles_world_x_64) { set_error("bssm_shard_peer_attach: null argument"); return BSSM_ERR_BAD_ARG; }
  if (!ctx->peer_inbox) { set_error("bssm_shard_peer_attach: call bssm_shard_peer_export first"); return BSSM_ERR_BAD_ARG; }
  BSSM_CK(cudaSetDevice(ctx->device));
  const char* hs = (const char*)handles_world_x_64;
  for (int g = 0; g < ctx->shard_world; g++) {
    if (g == ctx->shard_rank) { ctx->peer_ptr[g] = ctx->peer_inbox; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, hs + (size_t)g * sizeof(h), sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("bssm_shard_peer_attach: cudaIpcOpenMemHandle(rank %d): %s", g, cudaGetErrorString(e));
      cudaGetLastError();
      for (int k = 0; k < g; k++) {
        if (ctx->peer_ptr[k] && ctx->peer_ptr[k] != ctx->peer_inbox) cudaIpcCloseMemHandle(ctx->peer_ptr[k]);
        ctx->peer_ptr[k] = nullptr;
      }
      return BSSM_ERR_CUDA;
    }
    ctx->peer_ptr[g] = p;
  }
  ctx->peer_seq = 1;
  ctx->peer_on = 1;
  return BSSM_OK;
}

int bssm_shard_peer_detach(bssm_ctx* ctx) {
  if (!ctx) return BSSM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  peer_release(ctx);
  return BSSM_OK;
}

int bssm_shard_peer_active(const bssm_ctx* ctx) { return ctx ? ctx->peer_on : 0; }

int bssm_shard_finalize(bssm_ctx* ctx) {
  if (!ctx) return BSSM_OK;
  if (ctx->peer_inbox) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); peer_release(ctx); }
  if (ctx->nccl_comm) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    g_nccl.comm_destroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  ctx->shard_rank = 0; ctx->shard_world = 1;
  return BSSM_OK;
}

int bssm_shard_partition(int n, int world, int rank, int64_t* goff_out, int* nloc_out) {
  if (n < 1 || world < 1 || rank < 0 || rank >= world || !goff_out || !nloc_out) { set_error("bssm_shard_partition: bad argument"); return BSSM_ERR_BAD_ARG; }
  long long g; int nl;
  shard_partition(n, world, rank, &g, &nl);
  *goff_out = g; *nloc_out = nl;
  return BSSM_OK;
}

// One bootstrap filter of cfg->num_particles (GLOBAL count) particles sharded over the group of
// bssm_shard_init.  Collective: every rank calls it with identical cfg / y / theta.  Every rank gets
// the full (identical) outputs.
int bssm_filter_run_sharded(bssm_ctx* ctx, const bssm_filter_config* cfg, const double* y, const double* theta,
                            double capacity_factor, bssm_filter_result* res, int* n_local_final) {
  if (!ctx || !cfg || !y || !theta || !res) { set_error("bssm_filter_run_sharded: null argument"); return BSSM_ERR_BAD_ARG; }
  if (cfg->num_filters != 1) { set_error("bssm_filter_run_sharded: one filter per call (num_filters = 1)"); return BSSM_ERR_BAD_ARG; }
  if (cfg->num_particles < 1 || cfg->num_obs < 0 || cfg->dy < 1 || cfg->dy > 4) { set_error("bssm_filter_run_sharded: bad sizes"); return BSSM_ERR_BAD_ARG; }
  if (cfg->noise || cfg->return_particles || res->ancestors_history) { set_error("bssm_filter_run_sharded: injected noise / histories are not available on the sharded path"); return BSSM_ERR_UNSUPPORTED; }
  BSSM_CK(cudaSetDevice(ctx->device));
  int d, nth, nc;
  BSSM_TRY(model_dims(ctx, cfg->model, &d, &nth, &nc));
  const int C = 1, T = cfg->num_obs;
  ShardRun sh;
  sh.rank = ctx->shard_rank; sh.world = ctx->shard_world; sh.n_glob = cfg->num_particles;
  shard_partition(sh.n_glob, sh.world, sh.rank, &sh.goff0, &sh.nloc0);
  if (!(capacity_factor >= 1.0)) capacity_factor = 1.5;
  {
    const double want = capacity_factor * ((double)sh.n_glob / sh.world) + 1024.0;
    sh.cap = (int)(want < (double)sh.n_glob ? want : (double)sh.n_glob);
    if (sh.world == 1) sh.cap = sh.n_glob;
  }
  FilterDev f;
  memset(&f, 0, sizeof(f));
  f.C = C; f.N = sh.cap; f.T = T; f.dy = cfg->dy; f.d = d; f.n_per = nullptr;
  f.theta_stride = nth + nc; f.seed = cfg->seed;
  if (cfg->carry_weights) { set_error("particle-sharded filter: carry_weights is served by the general kernels only"); return BSSM_ERR_UNSUPPORTED; }
  f.algorithm = cfg->algorithm; f.ralg = cfg->resample_algorithm; f.threshold = cfg->threshold;
  FilterLaunch L;
  L.model = cfg->model; L.precision = cfg->precision; L.resample_fn = cfg->resample_fn; L.exact = 0; L.hist = 0; L.T = T;
  L.engine = BSSM_ENGINE_STREAM;
  if (!stream_supported(ctx, f, L)) { set_error("sharded filter: bootstrap filter of a 1-D built-in model with stratified / systematic resampling only"); return BSSM_ERR_UNSUPPORTED; }
  // per-filter scalars and outputs (the particle arrays belong to the streaming engine)
  double* sd; int* si;
  BSSM_TRY(scratch(ctx, SL_F_SCAL_D, (size_t)C * 4, &sd));
  BSSM_TRY(scratch(ctx, SL_F_SCAL_I, (size_t)C * 6, &si));
  f.M = sd; f.S = sd + C; f.loglike = sd + 2 * C; f.cur_ess = sd + 3 * C;
  f.alive = si; f.resample = si + C; f.status = si + 2 * C; f.early_exit = si + 3 * C; f.n_resampled = si + 4 * C; f.cur = si + 5 * C;
  BSSM_TRY(scratch(ctx, SL_F_ESS, (size_t)C * (T + 1), &f.ess));
  BSSM_TRY(scratch(ctx, SL_F_SEST, (size_t)C * (T + 1) * d, &f.state_est));
  BSSM_TRY(scratch(ctx, SL_F_LLH, (size_t)C * (T ? T : 1), &f.loglike_history));
  double *d_theta, *d_y; int* d_obs = nullptr; unsigned int* ids;
  BSSM_TRY(scratch(ctx, SL_F_THETA, (size_t)C * f.theta_stride, &d_theta));
  BSSM_TRY(scratch(ctx, SL_F_Y, (size_t)(T ? T : 1) * cfg->dy, &d_y));
  BSSM_TRY(scratch(ctx, SL_F_IDS, (size_t)2 * C, &ids));
  cudaStream_t st = ctx->stream;
  BSSM_CK(cudaMemcpyAsync(d_theta, theta, sizeof(double) * C * f.theta_stride, cudaMemcpyHostToDevice, st));
  if (T) BSSM_CK(cudaMemcpyAsync(d_y, y, sizeof(double) * T * cfg->dy, cudaMemcpyHostToDevice, st));
  if (cfg->obs_times && T) {
    BSSM_TRY(scratch(ctx, SL_F_OBS, (size_t)T, &d_obs));
    BSSM_CK(cudaMemcpyAsync(d_obs, cfg->obs_times, sizeof(int) * T, cudaMemcpyHostToDevice, st));
  }
  const unsigned int h_ids[2] = {cfg->stream_base, cfg->run_id};
  BSSM_CK(cudaMemcpyAsync(ids, h_ids, sizeof(h_ids), cudaMemcpyHostToDevice, st));
  f.theta = d_theta; f.y = d_y; f.obs_times = d_obs; f.stream = ids; f.run_id = ids + C;
  BSSM_TRY(filter_reset(ctx, f, nullptr));
  BSSM_CK(cudaEventRecord(ctx->ev0, st));
  BSSM_TRY(stream_filter_enqueue(ctx, f, L, &sh));
  BSSM_CK(cudaEventRecord(ctx->ev1, st));
  const size_t T1 = (size_t)T + 1;
#define DL(dst, src, count, type) if (dst) BSSM_CK(cudaMemcpyAsync(dst, src, (count) * sizeof(type), cudaMemcpyDeviceToHost, st))
  DL(res->loglike, f.loglike, (size_t)C, double);
  DL(res->loglike_history, f.loglike_history, (size_t)C * T, double);
  DL(res->ess, f.ess, (size_t)C * T1, double);
  DL(res->state_est, f.state_est, (size_t)C * T1 * d, double);
  DL(res->status, f.status, (size_t)C, int);
  DL(res->early_exit, f.early_exit, (size_t)C, int);
  DL(res->n_resampled, f.n_resampled, (size_t)C, int);
#undef DL
  StSeg h_seg[2]; int h_res[2];
  StSeg* d_seg; int* d_res;
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 7, (size_t)2 * C, &d_seg));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 6, (size_t)2 * C, &d_res));
  BSSM_CK(cudaMemcpyAsync(h_seg, d_seg, sizeof(h_seg), cudaMemcpyDeviceToHost, st));
  BSSM_CK(cudaMemcpyAsync(h_res, d_res, sizeof(h_res), cudaMemcpyDeviceToHost, st));
  BSSM_CK(cudaStreamSynchronize(st));
  BSSM_CK(cudaEventElapsedTime(&res->kernel_ms, ctx->ev0, ctx->ev1));
  if (n_local_final) {
    const int pp = (T + 1) & 1;
    *n_local_final = h_res[pp] ? h_seg[pp].nnloc : h_seg[pp].nloc;
  }
  return BSSM_OK;
}

}  // extern "C"
