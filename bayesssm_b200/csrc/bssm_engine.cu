// bssm_engine.cu -- C ABI of libbayesssm_b200.so: context, resamplers, particle filters.
// (PMMH lives in bssm_pmmh.cu, the persistent bootstrap-filter kernel in bssm_fast.cu.)
// No CPU fallback anywhere: every entry point needs a CUDA device.
#include "bssm_engine.cuh"
#include "bssm_fast.cuh"

#include <stdarg.h>

namespace bssm {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int scratch_get(bssm_ctx* ctx, int slot, size_t bytes, void** out) {
  Scratch& s = ctx->scratch[slot];
  if (bytes == 0) bytes = 16;
  if (s.cap < bytes) {
    if (s.p) { BSSM_CK(cudaStreamSynchronize(ctx->stream)); BSSM_CK(cudaFree(s.p)); s.p = nullptr; s.cap = 0; }
    size_t cap = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&s.p, cap);
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu bytes failed: %s", cap, cudaGetErrorString(e));
      s.p = nullptr;
      return BSSM_ERR_CUDA;
    }
    s.cap = cap;
  }
  *out = s.p;
  return BSSM_OK;
}

int check_launch(bssm_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("kernel launch failed (%s): %s", what, cudaGetErrorString(e));
    return BSSM_ERR_CUDA;
  }
  return BSSM_OK;
}

// ---------------------------------------------------------------------------------------------
// cdf pipeline
// ---------------------------------------------------------------------------------------------
__global__ void k_zero_sum_check(const double* total, int nseg, int* status, const int* enable) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg || !seg_on(enable, s)) return;
  if (status && status[s] == 0 && total[s] == 0.0) status[s] = BSSM_ERR_ZERO_SUM;
}

template <typename Src, typename MakeNorm>
int resample_cdf(bssm_ctx* ctx, const Src& src, MakeNorm make_norm, const RsArgs& a) {
  const int ntiles = (a.n + RS_TILE - 1) / RS_TILE;
  double* part; TileRec* rec; double* cstart; int* used; double* total;
  BSSM_TRY(scratch(ctx, SL_RS_PART, (size_t)a.nseg * ntiles, &part));
  BSSM_TRY(scratch(ctx, SL_RS_REC, (size_t)a.nseg * ntiles, &rec));
  BSSM_TRY(scratch(ctx, SL_RS_CSTART, (size_t)a.nseg * ntiles, &cstart));
  BSSM_TRY(scratch(ctx, SL_RS_USED, (size_t)a.nseg * ntiles, &used));
  if (a.total) total = a.total; else BSSM_TRY(scratch(ctx, SL_RS_TOTAL, (size_t)a.nseg, &total));
  dim3 grid(a.nseg, ntiles);
  cudaStream_t st = ctx->stream;
  // large inputs: the tile totals are scanned once per pass (k_tile_prefix) instead of summed again by every tile;
  // BSSM_RS_PREFIX_TILES moves the threshold (tests)
  int prefix_from = RS_PREFIX_TILES;
  if (const char* e = getenv("BSSM_RS_PREFIX_TILES")) prefix_from = atoi(e);
  double* pref = nullptr;
  if (ntiles > prefix_from) BSSM_TRY(scratch(ctx, SL_RS_PREF, (size_t)a.nseg * ntiles, &pref));
  auto tile_prefix = [&]() -> int {
    if (!pref) return BSSM_OK;
    k_tile_prefix<<<a.nseg, 256, 0, st>>>(part, a.n, a.n_per, ntiles, pref, a.enable);
    BSSM_LAUNCH(ctx, "k_tile_prefix");
    return BSSM_OK;
  };
  k_tile_sums<Src><<<grid, RS_THREADS, 0, st>>>(src, a.n, a.n_per, ntiles, part, a.status, a.validate, a.enable);
  BSSM_LAUNCH(ctx, "k_tile_sums");
  BSSM_TRY(tile_prefix());
  if (!a.exact) {
    k_tile_scan<Src><<<grid, RS_THREADS, 0, st>>>(src, a.n, a.n_per, ntiles, part, a.cdf, a.cdf_stride, rec, 0, a.enable, pref);
    BSSM_LAUNCH(ctx, "k_tile_scan");
    return BSSM_OK;
  }
  // pass 1: exact sequential total of the raw weights (src/resampling.cpp:20,47)
  k_tile_scan<Src><<<grid, RS_THREADS, 0, st>>>(src, a.n, a.n_per, ntiles, part, nullptr, 0, rec, 1, a.enable, pref);
  BSSM_LAUNCH(ctx, "k_tile_scan");
  k_chain<Src><<<a.nseg, 32, 0, st>>>(src, a.n, a.n_per, ntiles, rec, cstart, used, total, nullptr, 0, nullptr, a.enable);
  BSSM_LAUNCH(ctx, "k_chain");
  if (a.status) {
    k_zero_sum_check<<<(a.nseg + 127) / 128, 128, 0, st>>>(total, a.nseg, a.status, a.enable);
    BSSM_LAUNCH(ctx, "k_zero_sum_check");
  }
  // pass 2: exact sequential cumsum of w / total (src/resampling.cpp:24-25,51-52)
  auto srcn = make_norm(total);
  typedef decltype(srcn) SrcN;
  k_tile_sums<SrcN><<<grid, RS_THREADS, 0, st>>>(srcn, a.n, a.n_per, ntiles, part, nullptr, 0, a.enable);
  BSSM_LAUNCH(ctx, "k_tile_sums");
  BSSM_TRY(tile_prefix());
  k_tile_scan<SrcN><<<grid, RS_THREADS, 0, st>>>(srcn, a.n, a.n_per, ntiles, part, nullptr, 0, rec, 1, a.enable, pref);
  BSSM_LAUNCH(ctx, "k_tile_scan");
  k_chain<SrcN><<<a.nseg, 32, 0, st>>>(srcn, a.n, a.n_per, ntiles, rec, cstart, used, part /* pass-2 total (~1) is not needed; must not overwrite `total` */,
                                       a.cdf, a.cdf_stride, a.n_serial, a.enable);
  BSSM_LAUNCH(ctx, "k_chain");
  k_tile_exact<SrcN><<<grid, RS_THREADS, 0, st>>>(srcn, a.n, a.n_per, ntiles, cstart, used, a.cdf, a.cdf_stride, a.enable);
  BSSM_LAUNCH(ctx, "k_tile_exact");
  return BSSM_OK;
}

static int resample_cdf_plain(bssm_ctx* ctx, const double* d_w, size_t w_stride, const RsArgs& a) {
  SrcPlain src{d_w, w_stride};
  return resample_cdf(ctx, src, [=](const double* total) { return SrcPlainNorm{d_w, w_stride, total}; }, a);
}

// ---------------------------------------------------------------------------------------------
// particle filter orchestration
// ---------------------------------------------------------------------------------------------

template <typename Real>
static int resample_stage(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, int obs, int aux_stage, double* cdf) {
  RsArgs a;
  a.nseg = f.C; a.n = f.N; a.n_per = f.n_per; a.enable = f.resample; a.cdf = cdf; a.cdf_stride = (size_t)f.N;
  a.status = nullptr; a.validate = 0; a.exact = L.exact; a.n_serial = nullptr; a.total = nullptr;
  const Real* lw = (const Real*)(aux_stage ? f.lw_aux : f.lw);
  const double *M = f.M, *S = f.S;
  size_t stride = (size_t)f.N;
  SrcLogW<Real> src{lw, stride, M, S};
  BSSM_TRY(resample_cdf(ctx, src, [=](const double* total) { return SrcLogWNorm<Real>{lw, stride, M, S, total}; }, a));
  USrcFilter us;
  us.buf = aux_stage ? f.noise.u_resample_aux : f.noise.u_resample;
  us.N = f.N; us.injected = f.noise.injected; us.seed = f.seed; us.run_id = f.run_id; us.stream = f.stream;
  us.tag = aux_stage ? TAG_RESAMP_AUX_U : TAG_RESAMP_U; us.obs = obs;
  dim3 grid(f.C, f.nblk);
  k_search_gather<Real><<<grid, FT_THREADS, 0, ctx->stream>>>(f, us, L.resample_fn, obs, aux_stage, cdf);
  BSSM_LAUNCH(ctx, "k_search_gather");
  k_flip<<<(f.C + 127) / 128, 128, 0, ctx->stream>>>(f);
  BSSM_LAUNCH(ctx, "k_flip");
  return BSSM_OK;
}

// model kernels are launched through handles so that built-in (compiled in) and NVRTC-compiled user
// models share one orchestration
template <typename Model, typename Real> static ModelKernels builtin_kernels() {
  ModelKernels k;
  k.init = (void*)k_init<Model, Real>; k.weight = (void*)k_weight<Model, Real>; k.post = (void*)k_post<Model, Real>;
  k.has_aux = Model::HAS_AUX; k.has_move = Model::HAS_MOVE;
  return k;
}
static int launch_init(bssm_ctx* ctx, const ModelKernels& K, dim3 grid, FilterDev& f) {
  void* args[] = {&f};
  BSSM_CK(cudaLaunchKernel(K.init, grid, dim3(FT_THREADS), args, 0, ctx->stream));
  BSSM_LAUNCH(ctx, "k_init");
  return BSSM_OK;
}
static int launch_weight(bssm_ctx* ctx, const ModelKernels& K, dim3 grid, FilterDev& f, int obs, int flags, int wkind) {
  void* args[] = {&f, &obs, &flags, &wkind};
  BSSM_CK(cudaLaunchKernel(K.weight, grid, dim3(FT_THREADS), args, 0, ctx->stream));
  BSSM_LAUNCH(ctx, "k_weight");
  return BSSM_OK;
}
static int launch_post(bssm_ctx* ctx, const ModelKernels& K, dim3 grid, FilterDev& f, int obs) {
  void* args[] = {&f, &obs};
  BSSM_CK(cudaLaunchKernel(K.post, grid, dim3(FT_THREADS), args, 0, ctx->stream));
  BSSM_LAUNCH(ctx, "k_post");
  return BSSM_OK;
}

// one history row (particles and weights of all filters at time `row`): computed into ring slot row % 2 on the compute stream
// once the copy of row - 2 has left that slot, then copied to the caller's [C][T+1][..] buffers on the copy stream (2-D copies:
// one line per filter).  The destination is pinned for the duration of the call when the driver allows it (bssm_filter_run).
template <typename Real>
static int hist_row_out(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, dim3 grid, int row) {
  cudaStream_t st = ctx->stream;
  const int slot = row % f.hist_rows;
  if (!ctx->copy_stream) {
    BSSM_CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) { BSSM_CK(cudaEventCreateWithFlags(&ctx->ev_row[i], cudaEventDisableTiming)); BSSM_CK(cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming)); }
  }
  if (row >= f.hist_rows) BSSM_CK(cudaStreamWaitEvent(st, ctx->ev_free[slot], 0));
  k_history<Real><<<grid, FT_THREADS, 0, st>>>(f, row);
  BSSM_LAUNCH(ctx, "k_history");
  BSSM_CK(cudaEventRecord(ctx->ev_row[slot], st));
  BSSM_CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_row[slot], 0));
  const size_t T1 = (size_t)f.T + 1, N = (size_t)f.N, d = (size_t)f.d, C = (size_t)f.C;
  if (L.h_particles_history)
    BSSM_CK(cudaMemcpy2DAsync(L.h_particles_history + (size_t)row * d * N, T1 * d * N * sizeof(double),
                              f.particles_history + (size_t)slot * d * N, (size_t)f.hist_rows * d * N * sizeof(double),
                              d * N * sizeof(double), C, cudaMemcpyDeviceToHost, ctx->copy_stream));
  if (L.h_weights_history)
    BSSM_CK(cudaMemcpy2DAsync(L.h_weights_history + (size_t)row * N, T1 * N * sizeof(double),
                              f.weights_history + (size_t)slot * N, (size_t)f.hist_rows * N * sizeof(double),
                              N * sizeof(double), C, cudaMemcpyDeviceToHost, ctx->copy_stream));
  BSSM_CK(cudaEventRecord(ctx->ev_free[slot], ctx->copy_stream));
  return BSSM_OK;
}

template <typename Real>
static int run_filter_steps(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, double* cdf, const ModelKernels& K) {
  dim3 grid(f.C, f.nblk);
  cudaStream_t st = ctx->stream;
  if (f.algorithm == BSSM_APF && !K.has_aux) { set_error("model has no aux_log_likelihood_fn"); return BSSM_ERR_UNSUPPORTED; }
  if (f.algorithm == BSSM_RMPF && !K.has_move) { set_error("model has no move_fn"); return BSSM_ERR_UNSUPPORTED; }
  BSSM_TRY(launch_init(ctx, K, grid, f));
  k_finalize<sizeof(Real) == 4><<<f.C, 128, 0, st>>>(f, 0, 0);
  BSSM_LAUNCH(ctx, "k_finalize");
  if (L.hist) BSSM_TRY(hist_row_out<Real>(ctx, f, L, grid, 0));
  const bool may_resample = (f.algorithm == BSSM_RMPF) || (f.ralg != BSSM_SIS);
  for (int obs = 0; obs < L.T; obs++) {
    if (f.algorithm == BSSM_APF) {
      BSSM_TRY(launch_weight(ctx, K, grid, f, obs, WF_GAP, 1));
      k_finalize<sizeof(Real) == 4><<<f.C, 128, 0, st>>>(f, obs, 2);
      BSSM_LAUNCH(ctx, "k_finalize");
      BSSM_TRY(resample_stage<Real>(ctx, f, L, obs, 1, cdf));
      BSSM_TRY(launch_weight(ctx, K, grid, f, obs, WF_SECOND, 2));
    } else {
      BSSM_TRY(launch_weight(ctx, K, grid, f, obs, WF_GAP, 0));
    }
    k_finalize<sizeof(Real) == 4><<<f.C, 128, 0, st>>>(f, obs, 1);
    BSSM_LAUNCH(ctx, "k_finalize");
    if (may_resample) {
      BSSM_TRY(resample_stage<Real>(ctx, f, L, obs, 0, cdf));
      BSSM_TRY(launch_post(ctx, K, grid, f, obs));
      k_finalize<sizeof(Real) == 4><<<f.C, 128, 0, st>>>(f, obs, 3);
      BSSM_LAUNCH(ctx, "k_finalize");
    }
    if (L.hist) BSSM_TRY(hist_row_out<Real>(ctx, f, L, grid, obs + 1));
  }
  return BSSM_OK;
}

template <typename Model>
static int run_filter_model(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, double* cdf) {
  if (L.precision == BSSM_F64) return run_filter_steps<double>(ctx, f, L, cdf, builtin_kernels<Model, double>());
  return run_filter_steps<float>(ctx, f, L, cdf, builtin_kernels<Model, float>());
}

int model_dims(bssm_ctx* ctx, int model, int* d, int* ntheta, int* nconst) {
  if (model >= BSSM_USER_MODEL_BASE) {
    const UserModelInfo* u = user_model(ctx, model);
    if (!u) { set_error("unknown user model id %d", model); return BSSM_ERR_BAD_ARG; }
    *d = u->dims[0]; *ntheta = u->dims[1]; *nconst = u->dims[2];
    return BSSM_OK;
  }
#define MD(M) { *d = M::D; *ntheta = M::NTHETA; *nconst = M::NCONST; return BSSM_OK; }
  switch (model) {
    case BSSM_MODEL_AR_SIN: MD(ModelArSin)
    case BSSM_MODEL_LG: MD(ModelLG)
    case BSSM_MODEL_RW_DRIFT: MD(ModelRwDrift)
    case BSSM_MODEL_SIR_CB: MD(ModelSirCB)
    case BSSM_MODEL_AR_COS: MD(ModelArCos)
    case BSSM_MODEL_RW2D: MD(ModelRw2D)
    case BSSM_MODEL_SIR_GILLESPIE: MD(ModelSirGillespie)
  }
#undef MD
  set_error("unknown model id %d", model);
  return BSSM_ERR_BAD_ARG;
}

// Which engine serves a run.  The general kernels serve everything; the two bootstrap-filter engines serve the
// throughput precision (and BSSM_F64 when asked for by name):
//   persistent kernel  particles in registers, one launch per filter: single filters up to ~2 M particles and
//                      small batches -- latency-bound sizes
//   streaming engine   particles in HBM, two launches per observation: large batches [chains x particles] and
//                      single filters beyond the persistent kernel's capacity (measured: 1024 x 65536 runs
//                      1.3x faster here, N = 2^26 is only possible here)
// Returns -1 when an engine requested by name cannot serve the configuration.
int resolve_engine(bssm_ctx* ctx, const FilterDev& f_in, const FilterLaunch& L, bool injected, bool want_anc) {
  FilterDev f = f_in;
  f.noise.injected = injected ? 1 : 0;
  f.anc_history = want_anc ? (int*)1 : nullptr;
  if (L.engine == BSSM_ENGINE_GENERAL) return BSSM_ENGINE_GENERAL;
  if (L.engine == BSSM_ENGINE_PERSISTENT) return fast_supported(f, L) ? BSSM_ENGINE_PERSISTENT : -1;
  if (L.engine == BSSM_ENGINE_STREAM) return stream_supported(ctx, f, L) ? BSSM_ENGINE_STREAM : -1;
  if (L.precision != BSSM_F32) return BSSM_ENGINE_GENERAL;
  // the persistent kernel needs the CTAs of a filter co-resident: at most one full slice per SM
  const bool fast_ok = fast_supported(f, L) && (long long)f.N <= (long long)ctx->prop.multiProcessorCount * FAST_MAX_NB;
  const bool stream_ok = stream_supported(ctx, f, L);
  const long long total = (long long)f.C * f.N;
  // measured on B200 (scripts/bench_engines.py, G particle-timesteps/s streaming vs persistent): 1024 x 65536 141 vs 86,
  // 256 x 65536 103 vs 84, 128 x 65536 89 vs 82, 32 x 2^18 89 vs 75, 3 x 2^20 56 vs 48; but 64 x 65536 67 vs 77,
  // 8 x 2^19 64 vs 67, 2 x 2^20 47 vs 50, 1024 x 1000 27 vs 67: smaller batches and slices belong to the persistent kernel
  const bool stream_wins = (f.N >= (1 << 20) && f.C >= 3) || (f.N >= 32768 && total >= (1LL << 23));
  if (stream_ok && (!fast_ok || stream_wins)) return BSSM_ENGINE_STREAM;
  if (fast_ok) return BSSM_ENGINE_PERSISTENT;
  return BSSM_ENGINE_GENERAL;
}

// particle / weight / cdf arrays of the general kernels (f.nblk set)
static int general_arrays(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, bool need_aux, double** cdf_out) {
  const size_t C = f.C, N = f.N, d = f.d;
  const size_t rs = L.precision == BSSM_F64 ? 8 : 4;
  BSSM_TRY(scratch_get(ctx, SL_F_XA, C * d * N * rs, &f.xa));
  BSSM_TRY(scratch_get(ctx, SL_F_XB, C * d * N * rs, &f.xb));
  BSSM_TRY(scratch_get(ctx, SL_F_LW, C * N * rs, &f.lw));
  if (need_aux) {
    BSSM_TRY(scratch_get(ctx, SL_F_LWAUX, C * N * rs, &f.lw_aux));
    BSSM_TRY(scratch_get(ctx, SL_F_AUXG, C * N * rs, &f.auxg));
  }
  BSSM_TRY(scratch(ctx, SL_F_PART, C * f.nblk * PART_W, &f.part));
  BSSM_TRY(scratch(ctx, SL_F_CDF, C * N, cdf_out));
  return BSSM_OK;
}

// enqueue one batched filter run on ctx->stream (no synchronisation)
int filter_enqueue(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, double* cdf) {
  const int eng = resolve_engine(ctx, f, L, f.noise.injected != 0, f.anc_history != nullptr);
  if (eng < 0) {
    set_error("%s: configuration not supported by that engine (bootstrap filter of a 1-D built-in model, stratified / systematic resampling, no histories / injected noise)",
              L.engine == BSSM_ENGINE_PERSISTENT ? "BSSM_ENGINE_PERSISTENT" : "BSSM_ENGINE_STREAM");
    return BSSM_ERR_UNSUPPORTED;
  }
  if (eng == BSSM_ENGINE_PERSISTENT || eng == BSSM_ENGINE_STREAM) {
    const int st = eng == BSSM_ENGINE_PERSISTENT ? fast_filter_enqueue(ctx, f, L) : stream_filter_enqueue(ctx, f, L, nullptr);
    // A fast engine may still turn a configuration down when it sizes its launch (fewer resident blocks than the choice assumed:
    // a shared GPU, a part with fewer SMs) -- before anything has been launched.  Picked by AUTO, the general kernels then serve
    // the run; named by the caller, the refusal is the answer.
    if (st != BSSM_ERR_UNSUPPORTED || L.engine != BSSM_ENGINE_AUTO) return st;
    if (f.nblk < 1) { f.nblk = (int)(((size_t)f.N + FT_THREADS - 1) / FT_THREADS); if (f.nblk > 1024) f.nblk = 1024; }
    BSSM_TRY(general_arrays(ctx, f, L, false, &cdf));
  }
  if (!f.xa) { set_error("internal: general engine selected but its particle arrays were not set up"); return BSSM_ERR_BAD_ARG; }
  if (L.model >= BSSM_USER_MODEL_BASE) {   // NVRTC-compiled user model (bssm_nvrtc.cu)
    const UserModelInfo* u = user_model(ctx, L.model);
    if (!u) { set_error("unknown user model id %d", L.model); return BSSM_ERR_BAD_ARG; }
    if (L.precision == BSSM_F64) return run_filter_steps<double>(ctx, f, L, cdf, u->k64);
    return run_filter_steps<float>(ctx, f, L, cdf, u->k32);
  }
  switch (L.model) {
    case BSSM_MODEL_AR_SIN: return run_filter_model<ModelArSin>(ctx, f, L, cdf);
    case BSSM_MODEL_LG: return run_filter_model<ModelLG>(ctx, f, L, cdf);
    case BSSM_MODEL_RW_DRIFT: return run_filter_model<ModelRwDrift>(ctx, f, L, cdf);
    case BSSM_MODEL_SIR_CB: return run_filter_model<ModelSirCB>(ctx, f, L, cdf);
    case BSSM_MODEL_AR_COS: return run_filter_model<ModelArCos>(ctx, f, L, cdf);
    case BSSM_MODEL_RW2D: return run_filter_model<ModelRw2D>(ctx, f, L, cdf);
    case BSSM_MODEL_SIR_GILLESPIE: return run_filter_model<ModelSirGillespie>(ctx, f, L, cdf);
  }
  set_error("unknown model id %d", L.model);
  return BSSM_ERR_BAD_ARG;
}

__global__ void k_fill_ids(unsigned int* stream, unsigned int* run_id, int C, unsigned int stream_base, unsigned int rid) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { stream[c] = stream_base + (unsigned int)c; run_id[c] = rid; }
}
__global__ void k_fill_int(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_fill_uniform(double* p, size_t n, unsigned long long seed) {
  NoiseKey key = make_key(seed, 0u, 0u);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4x q = noise_quad(key, (unsigned int)(i >> 34), 0x55u, 0u, (unsigned int)(i >> 2));
    p[i] = word_to_unit_f64(q.w[i & 3]);
  }
}

// allocate the per-batch state of a filter run inside the context's scratch slots
int filter_setup(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, bool need_aux, bool want_anc, double** cdf_out, bool injected) {
  const size_t C = f.C, N = f.N, d = f.d, T = f.T;
  const size_t rs = L.precision == BSSM_F64 ? 8 : 4;
  f.nblk = (int)((N + FT_THREADS - 1) / FT_THREADS);
  if (f.nblk > 1024) f.nblk = 1024;
  if (f.nblk < 1) f.nblk = 1;
  // the particle / weight / cdf arrays belong to the general kernels; the persistent and streaming engines keep their own state
  const int eng = resolve_engine(ctx, f, L, injected, want_anc);
  const bool general = eng == BSSM_ENGINE_GENERAL || eng < 0;
  f.xa = f.xb = f.lw = f.lw_aux = f.auxg = nullptr; f.part = nullptr; *cdf_out = nullptr;
  if (general) BSSM_TRY(general_arrays(ctx, f, L, need_aux, cdf_out));
  double* sd; int* si;
  BSSM_TRY(scratch(ctx, SL_F_SCAL_D, C * 4, &sd));
  BSSM_TRY(scratch(ctx, SL_F_SCAL_I, C * 6, &si));
  f.M = sd; f.S = sd + C; f.loglike = sd + 2 * C; f.cur_ess = sd + 3 * C;
  f.alive = si; f.resample = si + C; f.status = si + 2 * C; f.early_exit = si + 3 * C; f.n_resampled = si + 4 * C; f.cur = si + 5 * C;
  BSSM_TRY(scratch(ctx, SL_F_ESS, C * (T + 1), &f.ess));
  BSSM_TRY(scratch(ctx, SL_F_SEST, C * (T + 1) * d, &f.state_est));
  BSSM_TRY(scratch(ctx, SL_F_LLH, C * (T ? T : 1), &f.loglike_history));
  if (L.hist) {
    // histories (R/particle_filter_core.R:100-116,242-264) never exist on the device as a whole -- (T + 1) N 16 bytes is 16.8 GB at
    // N = 2^20, T = 1000: a ring of two rows, each copied to the caller's buffers on a second stream while the next is computed
    f.hist_rows = 2;
    BSSM_TRY(scratch(ctx, SL_F_PH, C * f.hist_rows * d * N, &f.particles_history));
    BSSM_TRY(scratch(ctx, SL_F_WH, C * f.hist_rows * N, &f.weights_history));
  } else { f.particles_history = nullptr; f.weights_history = nullptr; f.hist_rows = 1; }
  if (want_anc) {
    BSSM_TRY(scratch(ctx, SL_F_ANC, C * (T ? T : 1) * N, &f.anc_history));
    BSSM_CK(cudaMemsetAsync(f.anc_history, 0, C * T * N * sizeof(int), ctx->stream));
    if (need_aux) {
      BSSM_TRY(scratch(ctx, SL_F_ANCA, C * (T ? T : 1) * N, &f.anc_aux_history));
      BSSM_CK(cudaMemsetAsync(f.anc_aux_history, 0, C * T * N * sizeof(int), ctx->stream));
    } else f.anc_aux_history = nullptr;
  } else { f.anc_history = nullptr; f.anc_aux_history = nullptr; }
  return BSSM_OK;
}

// reset the per-run state (outputs zero-filled as the reference's early-exit return, :189-202)
int filter_reset(bssm_ctx* ctx, FilterDev& f, const int* d_active) {
  const size_t C = f.C, T = f.T, d = f.d;
  cudaStream_t st = ctx->stream;
  BSSM_CK(cudaMemsetAsync(f.M, 0, C * 4 * sizeof(double), st));
  BSSM_CK(cudaMemsetAsync(f.alive, 0, C * 6 * sizeof(int), st));
  if (d_active) BSSM_CK(cudaMemcpyAsync(f.alive, d_active, C * sizeof(int), cudaMemcpyDeviceToDevice, st));
  else { k_fill_int<<<(int)((C + 255) / 256), 256, 0, st>>>(f.alive, (int)C, 1); BSSM_LAUNCH(ctx, "k_fill_int"); }
  BSSM_CK(cudaMemsetAsync(f.ess, 0, C * (T + 1) * sizeof(double), st));
  BSSM_CK(cudaMemsetAsync(f.state_est, 0, C * (T + 1) * d * sizeof(double), st));
  if (f.loglike_history) BSSM_CK(cudaMemsetAsync(f.loglike_history, 0, C * T * sizeof(double), st));
  return BSSM_OK;
}

}  // namespace bssm

using namespace bssm;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int bssm_abi_version(void) { return BSSM_ABI_VERSION; }
const char* bssm_last_error(void) { return g_err; }

int bssm_create(int device, bssm_ctx** out) {
  if (!out) { set_error("bssm_create: out is NULL"); return BSSM_ERR_BAD_ARG; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    set_error("no CUDA device available (%s); this engine has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return BSSM_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= ndev) { set_error("device %d out of range (0..%d)", device, ndev - 1); return BSSM_ERR_BAD_ARG; }
  BSSM_CK(cudaSetDevice(device));
  bssm_ctx* ctx = new bssm_ctx();
  ctx->device = device;
  BSSM_CK(cudaGetDeviceProperties(&ctx->prop, device));
  BSSM_CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  BSSM_CK(cudaEventCreate(&ctx->ev0));
  BSSM_CK(cudaEventCreate(&ctx->ev1));
  *out = ctx;
  return BSSM_OK;
}

void bssm_destroy(bssm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < SL_COUNT; i++) if (ctx->scratch[i].p) cudaFree(ctx->scratch[i].p);
  cudaEventDestroy(ctx->ev0);
  if (ctx->aux_stream) { cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->ev_mn_start); for (int i = 0; i < 2; i++) { cudaEventDestroy(ctx->ev_mn_ready[i]); cudaEventDestroy(ctx->ev_mn_free[i]); } }
  if (ctx->copy_stream) { cudaStreamDestroy(ctx->copy_stream); for (int i = 0; i < 2; i++) { cudaEventDestroy(ctx->ev_row[i]); cudaEventDestroy(ctx->ev_free[i]); } }
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int bssm_device_info(bssm_ctx* ctx, char* name, int* sm_count, int* cc_major, int* cc_minor, size_t* global_mem) {
  if (!ctx) { set_error("null context"); return BSSM_ERR_BAD_ARG; }
  if (name) { strncpy(name, ctx->prop.name, 255); name[255] = 0; }
  if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
  if (cc_major) *cc_major = ctx->prop.major;
  if (cc_minor) *cc_minor = ctx->prop.minor;
  if (global_mem) *global_mem = ctx->prop.totalGlobalMem;
  return BSSM_OK;
}
int64_t bssm_launch_count(bssm_ctx* ctx) { return ctx ? ctx->launches : 0; }
int bssm_synchronize(bssm_ctx* ctx) {
  BSSM_CK(cudaSetDevice(ctx->device));
  BSSM_CK(cudaStreamSynchronize(ctx->stream));
  return BSSM_OK;
}
int bssm_timer_start(bssm_ctx* ctx) { BSSM_CK(cudaEventRecord(ctx->ev0, ctx->stream)); return BSSM_OK; }
int bssm_timer_stop(bssm_ctx* ctx, float* ms_out) {
  BSSM_CK(cudaEventRecord(ctx->ev1, ctx->stream));
  BSSM_CK(cudaEventSynchronize(ctx->ev1));
  BSSM_CK(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
  return BSSM_OK;
}

int bssm_dev_alloc(bssm_ctx* ctx, size_t bytes, void** d_ptr) {
  BSSM_CK(cudaSetDevice(ctx->device));
  BSSM_CK(cudaMalloc(d_ptr, bytes ? bytes : 16));
  return BSSM_OK;
}
int bssm_dev_free(bssm_ctx* ctx, void* d_ptr) {
  BSSM_CK(cudaSetDevice(ctx->device));
  BSSM_CK(cudaStreamSynchronize(ctx->stream));
  BSSM_CK(cudaFree(d_ptr));
  return BSSM_OK;
}
int bssm_dev_upload(bssm_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  BSSM_CK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  BSSM_CK(cudaStreamSynchronize(ctx->stream));
  return BSSM_OK;
}
int bssm_dev_download(bssm_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  BSSM_CK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  BSSM_CK(cudaStreamSynchronize(ctx->stream));
  return BSSM_OK;
}
int bssm_dev_fill_uniform(bssm_ctx* ctx, double* d_ptr, size_t n, uint64_t seed) {
  k_fill_uniform<<<1184, 256, 0, ctx->stream>>>(d_ptr, n, seed);
  BSSM_LAUNCH(ctx, "k_fill_uniform");
  return BSSM_OK;
}

// ---- resamplers -------------------------------------------------------------------------------
int bssm_resample_device(bssm_ctx* ctx, int resample_fn, int batch, int n, const double* d_weights, const double* d_u,
                         int32_t* d_idx_out, int* d_status) {
  if (!ctx || batch <= 0 || n <= 0 || !d_weights || !d_u || !d_idx_out) { set_error("bssm_resample_device: bad argument"); return BSSM_ERR_BAD_ARG; }
  if (resample_fn < 0 || resample_fn > 2) { set_error("unknown resample_fn %d", resample_fn); return BSSM_ERR_BAD_ARG; }
  BSSM_CK(cudaSetDevice(ctx->device));
  double* cdf;
  BSSM_TRY(scratch(ctx, SL_API_CDF, (size_t)batch * n, &cdf));
  int* status = d_status;
  if (!status) BSSM_TRY(scratch(ctx, SL_RS_STATUS, (size_t)batch, &status));
  BSSM_CK(cudaMemsetAsync(status, 0, sizeof(int) * batch, ctx->stream));
  RsArgs a;
  a.nseg = batch; a.n = n; a.n_per = nullptr; a.enable = nullptr; a.cdf = cdf; a.cdf_stride = (size_t)n;
  a.status = status; a.validate = 1; a.exact = 1; a.n_serial = nullptr; a.total = nullptr;
  BSSM_TRY(resample_cdf_plain(ctx, d_weights, (size_t)n, a));
  USrcBuf us{d_u, resample_fn == BSSM_SYSTEMATIC ? (size_t)1 : (size_t)n};
  int gx = (n + 255) / 256; if (gx > 4096) gx = 4096;
  k_search<USrcBuf><<<dim3(batch, gx), 256, 0, ctx->stream>>>(us, resample_fn, n, nullptr, cdf, (size_t)n, d_idx_out, (size_t)n, nullptr);
  BSSM_LAUNCH(ctx, "k_search");
  return BSSM_OK;
}

static int resample_host(bssm_ctx* ctx, int fn, int n, const double* w, const double* u, int nu, int32_t* idx_out) {
  if (!ctx || n <= 0 || !w || !u || !idx_out) { set_error("resample: bad argument (n must be >= 1)"); return BSSM_ERR_BAD_ARG; }
  BSSM_CK(cudaSetDevice(ctx->device));
  double *dw, *du; int32_t* di; int* dst;
  BSSM_TRY(scratch(ctx, SL_API_W, (size_t)n, &dw));
  BSSM_TRY(scratch(ctx, SL_API_U, (size_t)nu, &du));
  BSSM_TRY(scratch(ctx, SL_API_IDX, (size_t)n, &di));
  BSSM_TRY(scratch(ctx, SL_RS_STATUS, (size_t)1, &dst));
  BSSM_CK(cudaMemcpyAsync(dw, w, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  BSSM_CK(cudaMemcpyAsync(du, u, sizeof(double) * nu, cudaMemcpyHostToDevice, ctx->stream));
  BSSM_TRY(bssm_resample_device(ctx, fn, 1, n, dw, du, di, dst));
  int status = 0;
  BSSM_CK(cudaMemcpyAsync(&status, dst, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  BSSM_CK(cudaMemcpyAsync(idx_out, di, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BSSM_CK(cudaStreamSynchronize(ctx->stream));
  if (status == BSSM_ERR_NEGATIVE_WEIGHT) { set_error("Weights must be non-negative"); return status; }
  if (status == BSSM_ERR_ZERO_SUM) { set_error("Sum of weights must be greater than 0"); return status; }
  if (status == BSSM_ERR_NAN_WEIGHT) { set_error("Weights must not be NaN"); return status; }
  return BSSM_OK;
}
int bssm_resample_stratified(bssm_ctx* ctx, int n, const double* weights, const double* u, int32_t* idx_out) {
  return resample_host(ctx, BSSM_STRATIFIED, n, weights, u, n, idx_out);
}
int bssm_resample_systematic(bssm_ctx* ctx, int n, const double* weights, double u, int32_t* idx_out) {
  return resample_host(ctx, BSSM_SYSTEMATIC, n, weights, &u, 1, idx_out);
}
int bssm_resample_multinomial(bssm_ctx* ctx, int n, const double* weights, const double* u, int32_t* idx_out) {
  return resample_host(ctx, BSSM_MULTINOMIAL, n, weights, u, n, idx_out);
}

int bssm_resample_cdf(bssm_ctx* ctx, int n, const double* weights, double* cdf_out, double* total_out, int64_t* n_serial_out) {
  if (!ctx || n <= 0 || !weights) { set_error("bssm_resample_cdf: bad argument"); return BSSM_ERR_BAD_ARG; }
  BSSM_CK(cudaSetDevice(ctx->device));
  double *dw, *cdf, *total; int* dst; long long* dser;
  BSSM_TRY(scratch(ctx, SL_API_W, (size_t)n, &dw));
  BSSM_TRY(scratch(ctx, SL_API_CDF, (size_t)n, &cdf));
  BSSM_TRY(scratch(ctx, SL_RS_STATUS, (size_t)1, &dst));
  BSSM_TRY(scratch(ctx, SL_RS_TOTAL, (size_t)1, &total));
  BSSM_TRY(scratch(ctx, SL_RS_NSERIAL, (size_t)1, &dser));
  BSSM_CK(cudaMemcpyAsync(dw, weights, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  BSSM_CK(cudaMemsetAsync(dst, 0, sizeof(int), ctx->stream));
  RsArgs a;
  a.nseg = 1; a.n = n; a.n_per = nullptr; a.enable = nullptr; a.cdf = cdf; a.cdf_stride = (size_t)n;
  a.status = dst; a.validate = 1; a.exact = 1; a.n_serial = dser; a.total = total;
  BSSM_TRY(resample_cdf_plain(ctx, dw, (size_t)n, a));
  int status = 0; long long ser = 0;
  BSSM_CK(cudaMemcpyAsync(&status, dst, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  BSSM_CK(cudaMemcpyAsync(&ser, dser, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (cdf_out) BSSM_CK(cudaMemcpyAsync(cdf_out, cdf, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (total_out) BSSM_CK(cudaMemcpyAsync(total_out, total, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BSSM_CK(cudaStreamSynchronize(ctx->stream));
  if (n_serial_out) *n_serial_out = ser;
  if (status) { set_error("invalid weights (status %d)", status); return status; }
  return BSSM_OK;
}

// ---- particle filters -------------------------------------------------------------------------
int bssm_model_dims(bssm_ctx* ctx, int model, int* d, int* ntheta, int* nconst) {
  (void)ctx;
  int dd, nt, nc;
  BSSM_TRY(model_dims(ctx, model, &dd, &nt, &nc));
  if (d) *d = dd;
  if (ntheta) *ntheta = nt;
  if (nconst) *nconst = nc;
  return BSSM_OK;
}

static int upload_opt(bssm_ctx* ctx, int slot, const double* h, size_t count, const double** d_out) {
  *d_out = nullptr;
  if (!h || !count) return BSSM_OK;
  double* d;
  BSSM_TRY(scratch(ctx, slot, count, &d));
  BSSM_CK(cudaMemcpyAsync(d, h, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  *d_out = d;
  return BSSM_OK;
}

static int filter_validate(const bssm_filter_config* cfg) {
  if (cfg->num_particles < 1) { set_error("Assertion on 'num_particles' failed: must be >= 1"); return BSSM_ERR_BAD_ARG; }
  if (cfg->num_obs < 0 || cfg->dy < 1 || cfg->dy > 4) { set_error("bad y dimensions (T=%d, dy=%d)", cfg->num_obs, cfg->dy); return BSSM_ERR_BAD_ARG; }
  if (cfg->num_filters < 1) { set_error("num_filters must be >= 1"); return BSSM_ERR_BAD_ARG; }
  if (cfg->algorithm < 0 || cfg->algorithm > 2 || cfg->resample_algorithm < 0 || cfg->resample_algorithm > 2 ||
      cfg->resample_fn < 0 || cfg->resample_fn > 2) { set_error("unknown algorithm / resample_algorithm / resample_fn"); return BSSM_ERR_BAD_ARG; }
  if (cfg->model == BSSM_MODEL_SIR_GILLESPIE && cfg->noise) {
    set_error("model %d draws a data-dependent number of uniforms per transition: injected noise buffers are not supported", cfg->model);
    return BSSM_ERR_UNSUPPORTED;
  }
  if (cfg->obs_times) {
    int prev = 1;   // R/particle_filter_core.R:69: checkmate::assert_integerish(obs_times, lower = 1, sorted = TRUE)
    for (int i = 0; i < cfg->num_obs; i++) {
      if (cfg->obs_times[i] < prev) { set_error("obs_times must be non-decreasing positive integers"); return BSSM_ERR_BAD_ARG; }
      prev = cfg->obs_times[i];
    }
  }
  return BSSM_OK;
}

int bssm_filter_run(bssm_ctx* ctx, const bssm_filter_config* cfg, const double* y, const double* theta, bssm_filter_result* res) {
  if (!ctx || !cfg || !y || !theta || !res) { set_error("bssm_filter_run: null argument"); return BSSM_ERR_BAD_ARG; }
  BSSM_TRY(filter_validate(cfg));
  BSSM_CK(cudaSetDevice(ctx->device));
  int d, nth, nc;
  BSSM_TRY(model_dims(ctx, cfg->model, &d, &nth, &nc));
  if (cfg->noise && cfg->model >= BSSM_USER_MODEL_BASE) {
    const UserModelInfo* um = user_model(ctx, cfg->model);
    if (um && um->dims[12]) { set_error("user model %d draws its transition uniforms on demand (DYN_U): injected noise buffers are not supported", cfg->model); return BSSM_ERR_UNSUPPORTED; }
  }
  const int C = cfg->num_filters, N = cfg->num_particles, T = cfg->num_obs;
  FilterDev f;
  memset(&f, 0, sizeof(f));
  f.C = C; f.N = N; f.T = T; f.dy = cfg->dy; f.d = d; f.n_per = nullptr;
  f.theta_stride = nth + nc; f.seed = cfg->seed;
  f.algorithm = cfg->algorithm;
  f.ralg = cfg->algorithm == BSSM_RMPF ? BSSM_SISR : cfg->resample_algorithm;
  f.threshold = cfg->threshold;
  f.carry = cfg->carry_weights ? 1 : 0;
  if (f.carry && cfg->algorithm == BSSM_APF) { set_error("carry_weights: not defined for the auxiliary filter's two-stage weights"); return BSSM_ERR_UNSUPPORTED; }
  FilterLaunch L;
  L.model = cfg->model; L.precision = cfg->precision; L.resample_fn = cfg->resample_fn;
  L.exact = cfg->exact_resampling < 0 ? (cfg->precision == BSSM_F64) : cfg->exact_resampling;
  L.hist = cfg->return_particles; L.T = T; L.engine = cfg->engine;
  // the caller's history buffers stay pinned for exactly this call, whichever way it ends (an error return included: the guard
  // drains both streams first, so no copy is in flight into memory it un-pins)
  struct HostPins {
    bssm_ctx* ctx; void* p[2] = {nullptr, nullptr};
    ~HostPins() {
      if (!p[0] && !p[1]) return;
      cudaStreamSynchronize(ctx->stream);
      if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
      for (void* q : p) if (q) cudaHostUnregister(q);
      cudaGetLastError();
    }
  } pins{ctx};
  if (cfg->return_particles) {
    // rows stream to the caller's buffers while the filter runs: pin them for the call (a pageable destination still works, the
    // copies then stage through the driver's own pinned buffer and the copy stream runs behind)
    L.h_particles_history = res->particles_history; L.h_weights_history = res->weights_history;
    const size_t T1h = (size_t)T + 1;
    if (res->particles_history && cudaHostRegister(res->particles_history, (size_t)C * T1h * d * N * sizeof(double), cudaHostRegisterDefault) == cudaSuccess) pins.p[0] = res->particles_history;
    if (res->weights_history && cudaHostRegister(res->weights_history, (size_t)C * T1h * N * sizeof(double), cudaHostRegisterDefault) == cudaSuccess) pins.p[1] = res->weights_history;
    cudaGetLastError();   // a refused registration is not an error of the call
  }
  const bool need_aux = cfg->algorithm == BSSM_APF;
  const bool want_anc = res->ancestors_history != nullptr || res->ancestors_aux_history != nullptr;
  double* cdf;
  BSSM_TRY(filter_setup(ctx, f, L, need_aux, want_anc, &cdf, cfg->noise != nullptr));
  // inputs
  double* d_theta; double* d_y; int* d_obs = nullptr; unsigned int* ids;
  BSSM_TRY(scratch(ctx, SL_F_THETA, (size_t)C * f.theta_stride, &d_theta));
  BSSM_TRY(scratch(ctx, SL_F_Y, (size_t)(T ? T : 1) * cfg->dy, &d_y));
  BSSM_TRY(scratch(ctx, SL_F_IDS, (size_t)2 * C, &ids));
  BSSM_CK(cudaMemcpyAsync(d_theta, theta, sizeof(double) * C * f.theta_stride, cudaMemcpyHostToDevice, ctx->stream));
  if (T) BSSM_CK(cudaMemcpyAsync(d_y, y, sizeof(double) * T * cfg->dy, cudaMemcpyHostToDevice, ctx->stream));
  if (cfg->obs_times && T) {
    BSSM_TRY(scratch(ctx, SL_F_OBS, (size_t)T, &d_obs));
    BSSM_CK(cudaMemcpyAsync(d_obs, cfg->obs_times, sizeof(int) * T, cudaMemcpyHostToDevice, ctx->stream));
  }
  f.theta = d_theta; f.y = d_y; f.obs_times = d_obs; f.stream = ids; f.run_id = ids + C;
  k_fill_ids<<<(C + 127) / 128, 128, 0, ctx->stream>>>(ids, ids + C, C, cfg->stream_base, cfg->run_id);
  BSSM_LAUNCH(ctx, "k_fill_ids");
  // injected noise
  memset(&f.noise, 0, sizeof(f.noise));
  if (cfg->noise) {
    const bssm_noise_buffers* nb = cfg->noise;
    int n_time = T ? (cfg->obs_times ? cfg->obs_times[T - 1] : T) : 0;
    const size_t NN = (size_t)N;
    // slot counts come from the model; the caller sized the buffers with bssm_model_noise_dims
    int nzi, nui, nzt, nut, nzm, num;
    BSSM_TRY(bssm_model_noise_dims(ctx, cfg->model, &nzi, &nui, &nzt, &nut, &nzm, &num));
    f.noise.injected = 1;
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 0, nb->z_init, nzi * NN, &f.noise.z_init));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 1, nb->u_init, nui * NN, &f.noise.u_init));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 2, nb->z_trans, (size_t)n_time * nzt * NN, &f.noise.z_trans));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 3, nb->u_trans, (size_t)n_time * nut * NN, &f.noise.u_trans));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 4, nb->z_trans2, (size_t)T * nzt * NN, &f.noise.z_trans2));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 5, nb->u_trans2, (size_t)T * nut * NN, &f.noise.u_trans2));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 6, nb->u_resample, (size_t)T * NN, &f.noise.u_resample));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 7, nb->u_resample_aux, (size_t)T * NN, &f.noise.u_resample_aux));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 8, nb->z_move, (size_t)T * nzm * NN, &f.noise.z_move));
    BSSM_TRY(upload_opt(ctx, SL_F_NOISE0 + 9, nb->u_move, (size_t)T * num * NN, &f.noise.u_move));
    bool ok = (nzi == 0 || f.noise.z_init) && (nui == 0 || f.noise.u_init) && (nzt == 0 || f.noise.z_trans) &&
              (nut == 0 || f.noise.u_trans);
    if (cfg->algorithm == BSSM_APF) ok = ok && (nzt == 0 || f.noise.z_trans2) && (nut == 0 || f.noise.u_trans2) && f.noise.u_resample_aux;
    if (cfg->algorithm == BSSM_RMPF) ok = ok && (nzm == 0 || f.noise.z_move) && (num == 0 || f.noise.u_move);
    if (f.ralg != BSSM_SIS) ok = ok && f.noise.u_resample;
    if (!ok) { set_error("injected noise: a buffer this model/algorithm consumes is NULL"); return BSSM_ERR_BAD_ARG; }
  }
  BSSM_TRY(filter_reset(ctx, f, nullptr));
  BSSM_CK(cudaEventRecord(ctx->ev0, ctx->stream));
  BSSM_TRY(filter_enqueue(ctx, f, L, cdf));
  BSSM_CK(cudaEventRecord(ctx->ev1, ctx->stream));
  // outputs
  cudaStream_t st = ctx->stream;
  const size_t T1 = (size_t)T + 1;
#define DL(dst, src, count, type) if (dst) BSSM_CK(cudaMemcpyAsync(dst, src, (count) * sizeof(type), cudaMemcpyDeviceToHost, st))
  DL(res->loglike, f.loglike, (size_t)C, double);
  DL(res->loglike_history, f.loglike_history, (size_t)C * T, double);
  DL(res->ess, f.ess, (size_t)C * T1, double);
  DL(res->state_est, f.state_est, (size_t)C * T1 * d, double);
  DL(res->status, f.status, (size_t)C, int);
  DL(res->early_exit, f.early_exit, (size_t)C, int);
  DL(res->n_resampled, f.n_resampled, (size_t)C, int);
  if (want_anc) {
    DL(res->ancestors_history, f.anc_history, (size_t)C * T * N, int);
    if (need_aux) DL(res->ancestors_aux_history, f.anc_aux_history, (size_t)C * T * N, int);
  }
#undef DL
  cudaError_t e_sync = cudaStreamSynchronize(st);
  if (e_sync == cudaSuccess && cfg->return_particles && ctx->copy_stream) e_sync = cudaStreamSynchronize(ctx->copy_stream);
  BSSM_CK(e_sync);
  BSSM_CK(cudaEventElapsedTime(&res->kernel_ms, ctx->ev0, ctx->ev1));
  return BSSM_OK;
}

// device-resident variant: y [T][dy] and theta [C][ntheta+nconst] already in HBM, Philox noise, no
// histories; only the log-likelihoods are written (device).  No host<->device traffic.
int bssm_filter_run_device(bssm_ctx* ctx, const bssm_filter_config* cfg, const double* d_y, const double* d_theta,
                           double* d_loglike, float* kernel_ms) {
  if (!ctx || !cfg || !d_y || !d_theta) { set_error("bssm_filter_run_device: null argument"); return BSSM_ERR_BAD_ARG; }
  BSSM_TRY(filter_validate(cfg));
  if (cfg->noise || cfg->return_particles) { set_error("bssm_filter_run_device: injected noise / histories need bssm_filter_run"); return BSSM_ERR_UNSUPPORTED; }
  BSSM_CK(cudaSetDevice(ctx->device));
  int d, nth, nc;
  BSSM_TRY(model_dims(ctx, cfg->model, &d, &nth, &nc));
  const int C = cfg->num_filters, T = cfg->num_obs;
  FilterDev f;
  memset(&f, 0, sizeof(f));
  f.C = C; f.N = cfg->num_particles; f.T = T; f.dy = cfg->dy; f.d = d; f.theta_stride = nth + nc; f.seed = cfg->seed;
  f.algorithm = cfg->algorithm;
  f.ralg = cfg->algorithm == BSSM_RMPF ? BSSM_SISR : cfg->resample_algorithm;
  f.threshold = cfg->threshold;
  f.carry = cfg->carry_weights ? 1 : 0;
  if (f.carry && cfg->algorithm == BSSM_APF) { set_error("carry_weights: not defined for the auxiliary filter's two-stage weights"); return BSSM_ERR_UNSUPPORTED; }
  FilterLaunch L;
  L.model = cfg->model; L.precision = cfg->precision; L.resample_fn = cfg->resample_fn;
  L.exact = cfg->exact_resampling < 0 ? (cfg->precision == BSSM_F64) : cfg->exact_resampling;
  L.hist = 0; L.T = T; L.engine = cfg->engine;
  double* cdf;
  BSSM_TRY(filter_setup(ctx, f, L, cfg->algorithm == BSSM_APF, false, &cdf));
  unsigned int* ids; int* d_obs = nullptr;
  BSSM_TRY(scratch(ctx, SL_F_IDS, (size_t)2 * C, &ids));
  if (cfg->obs_times && T) {
    BSSM_TRY(scratch(ctx, SL_F_OBS, (size_t)T, &d_obs));
    BSSM_CK(cudaMemcpyAsync(d_obs, cfg->obs_times, sizeof(int) * T, cudaMemcpyHostToDevice, ctx->stream));
  }
  f.theta = d_theta; f.y = d_y; f.obs_times = d_obs; f.stream = ids; f.run_id = ids + C;
  k_fill_ids<<<(C + 127) / 128, 128, 0, ctx->stream>>>(ids, ids + C, C, cfg->stream_base, cfg->run_id);
  BSSM_LAUNCH(ctx, "k_fill_ids");
  BSSM_TRY(filter_reset(ctx, f, nullptr));
  BSSM_CK(cudaEventRecord(ctx->ev0, ctx->stream));
  BSSM_TRY(filter_enqueue(ctx, f, L, cdf));
  BSSM_CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (d_loglike) BSSM_CK(cudaMemcpyAsync(d_loglike, f.loglike, sizeof(double) * C, cudaMemcpyDeviceToDevice, ctx->stream));
  if (kernel_ms) {
    BSSM_CK(cudaEventSynchronize(ctx->ev1));
    BSSM_CK(cudaEventElapsedTime(kernel_ms, ctx->ev0, ctx->ev1));
  }
  return BSSM_OK;
}

int bssm_model_noise_dims(bssm_ctx* ctx, int model, int* nz_init, int* nu_init, int* nz_trans, int* nu_trans, int* nz_move, int* nu_move) {
  if (model >= BSSM_USER_MODEL_BASE) {
    const UserModelInfo* u = user_model(ctx, model);
    if (!u) { set_error("unknown user model id %d", model); return BSSM_ERR_BAD_ARG; }
    *nz_init = u->dims[3]; *nu_init = u->dims[4]; *nz_trans = u->dims[5]; *nu_trans = u->dims[6]; *nz_move = u->dims[7]; *nu_move = u->dims[8];
    return BSSM_OK;
  }
#define ND(M) { *nz_init = M::NZ_INIT; *nu_init = M::NU_INIT; *nz_trans = M::NZ_TRANS; *nu_trans = M::NU_TRANS; *nz_move = M::NZ_MOVE; *nu_move = M::NU_MOVE; return BSSM_OK; }
  switch (model) {
    case BSSM_MODEL_AR_SIN: ND(ModelArSin)
    case BSSM_MODEL_LG: ND(ModelLG)
    case BSSM_MODEL_RW_DRIFT: ND(ModelRwDrift)
    case BSSM_MODEL_SIR_CB: ND(ModelSirCB)
    case BSSM_MODEL_AR_COS: ND(ModelArCos)
    case BSSM_MODEL_RW2D: ND(ModelRw2D)
    case BSSM_MODEL_SIR_GILLESPIE: ND(ModelSirGillespie)
  }
#undef ND
  set_error("unknown model id %d", model);
  return BSSM_ERR_BAD_ARG;
}

}  // extern "C"
