// bssm_stream.cu -- host side of the streaming bootstrap-filter engine (bssm_stream.cuh): two kernel
// launches per observation, all decisions on the device, no host synchronisation inside a filter.
// With a shard communicator (bssm_shard.cu) one ncclAllGather of a 64-byte record per filter and a
// one-thread merge kernel sit between the two kernels of an observation.
#include "bssm_engine.cuh"
#include "bssm_stream.cuh"

#include <math.h>
#include <stdlib.h>

namespace bssm {

bool stream_supported(bssm_ctx* ctx, const FilterDev& f, const FilterLaunch& L) {
  if (f.algorithm != BSSM_BPF || L.hist || f.noise.injected || f.anc_history || f.carry) return false;
  if (L.model >= BSSM_USER_MODEL_BASE) {   // NVRTC user model: its shape decides (bssm_nvrtc.cu)
    const UserModelInfo* u = user_model(ctx, L.model);
    return u && u->stream_ok;
  }
  if (!(L.model == BSSM_MODEL_AR_SIN || L.model == BSSM_MODEL_LG || L.model == BSSM_MODEL_AR_COS || L.model == BSSM_MODEL_RW_DRIFT)) return false;
  return true;
}

// kernels are launched through handles, so that built-in models (function addresses) and NVRTC-compiled user
// models (cudaKernel_t of the compiled library) share one orchestration
// pdl: programmatic dependent launch -- the kernel may become resident while its predecessor in the stream drains (both kernels
// call griddepcontrol.launch_dependents / .wait, bssm_stream.cuh); BSSM_ST_PDL=0 turns it off
static int st_launch(bssm_ctx* ctx, void* kern, dim3 grid, int threads, StreamParams& P, int* obs, const char* what, bool pdl) {
  void* args[] = {&P, obs};
  if (pdl) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    BSSM_CK(cudaLaunchKernelExC(&cfg, kern, args));
  } else {
    BSSM_CK(cudaLaunchKernel(kern, grid, dim3(threads), args, 0, ctx->stream));
  }
  BSSM_LAUNCH(ctx, what);
  return BSSM_OK;
}

template <typename Real, int PPT, int THREADS>
static int stream_launch(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, const ShardRun* sh, const StreamKernels& K) {
  constexpr int ST_THREADS = THREADS;
  constexpr int TS = ST_THREADS * PPT;
  cudaStream_t st = ctx->stream;
  StreamParams P;
  memset(&P, 0, sizeof(P));
  P.f = f; P.resample_fn = L.resample_fn;
  st_fill_round_keys(P);
  const int C = f.C;
  long long goff0 = 0; int nloc0 = f.N; int cap = f.N;
  if (sh) {
    P.sharded = 1; P.rank = sh->rank; P.world = sh->world; P.n_glob = sh->n_glob;
    goff0 = sh->goff0; nloc0 = sh->nloc0; cap = sh->cap;
  } else { P.rank = 0; P.world = 1; }
  P.cap = cap;
  P.log_n = f.n_per ? nan("") : log((double)(sh ? sh->n_glob : f.N));
  P.nt = (cap + 4 + TS - 1) / TS;
  P.xstride = (size_t)P.nt * TS;
  const size_t rs = sizeof(Real);
  // blocks per filter: the resident block slots of the chip split over the filters, each block walking a
  // contiguous range of tiles; among a few candidates take the one with the fewest rounds (waves x tiles
  // per block).  The three kernels share the block -> tile ranges, so the scarcer kernel sets the slots.
  {
    int ps = 0, pr = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, (const void*)K.step, ST_THREADS, 0) != cudaSuccess) { ps = 0; cudaGetLastError(); }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr, (const void*)K.resample, ST_THREADS, 0) != cudaSuccess) { pr = 0; cudaGetLastError(); }
    int per_sm = ps < pr ? ps : pr;
    if (per_sm < 1) per_sm = 1024 / ST_THREADS;   // the kernels are built for 1024 threads per SM (__launch_bounds__)
    const long long slots = (long long)per_sm * ctx->prop.multiProcessorCount;
    long long b0 = (slots + C - 1) / C;
    const long long bmin = 1;
    if (b0 < bmin) b0 = bmin;
    if (b0 > P.nt) b0 = P.nt;
    // cost in tile-times: the blocks' work spread over the resident slots plus one block's length as the tail;
    // a block pays ~2 tile-times of prologue / ticket on top of its tiles
    long long best = b0; double best_cost = -1.0;
    for (long long b = b0; b <= P.nt && b <= 8 * b0 + 8; b++) {
      const double len = (double)((P.nt + b - 1) / b) + 2.0;
      const double blocks = (double)C * (double)b;
      const double cost = (blocks <= (double)slots ? len : blocks * len / (double)slots + len);
      if (best_cost < 0 || cost < best_cost) { best = b; best_cost = cost; }
    }
    if (const char* e = getenv("BSSM_ST_BPC")) { int v = atoi(e); if (v >= bmin && v <= P.nt) best = v; }
    P.bpc = (int)best;
  }
  const bool mn = L.resample_fn == BSSM_MULTINOMIAL;
  // Batches on one GPU: every observation inside one cooperative launch per group of filters (k_st_chain, bssm_stream.cuh) --
  // the blocks of a filter meet through two words instead of two kernel boundaries per observation.  BSSM_ST_CHAIN=0 / 1
  // turns it off / forces it (A/B runs); multinomial resampling and the sharded filter have kernels or a collective between
  // the two bodies and keep the launch-per-body form, as do NVRTC user models.
  // Measured (PMMH geometry, 65 536 particles per filter): +8 % at 256 filters, +6 % at 128, -8 % at 1024 -- a big batch keeps the
  // chip busy across kernel boundaries anyway, and the cooperative form gives up the block slots its groups cannot fill: it
  // takes batches of up to a quarter of the resident block slots (384 and 512 filters: level with the launch-per-body form).
  bool chain = K.chain != nullptr && !sh && !mn && L.T > 0 && ctx->prop.cooperativeLaunch;
  int chain_slots = 0;
  if (chain) {
    int pk = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pk, (const void*)K.chain, ST_THREADS, 0) != cudaSuccess) { pk = 0; cudaGetLastError(); }
    chain_slots = pk * ctx->prop.multiProcessorCount;
    if (chain_slots < 1) chain = false;
  }
  if (chain) {
    const char* e = getenv("BSSM_ST_CHAIN");
    chain = e ? atoi(e) != 0 : (C >= 16 && 4 * C <= chain_slots);
  }
  if (chain) {
    // blocks per filter: the resident slots split over the filters of a group (all blocks of a group are co-resident)
    long long b = chain_slots / (C < chain_slots ? C : chain_slots);
    if (b > P.nt) b = P.nt;
    if (b < 1) b = 1;
    if (const char* e = getenv("BSSM_ST_BPC")) { int v = atoi(e); if (v >= 1 && v <= P.nt && v <= chain_slots) b = v; }
    P.bpc = (int)b;
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 12, (size_t)C, &P.epoch));
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 13, (size_t)C, &P.bar2));
  }
  BSSM_TRY(scratch_get(ctx, SL_ST_BASE + 0, (size_t)C * P.xstride * rs, &P.x0));
  BSSM_TRY(scratch_get(ctx, SL_ST_BASE + 1, (size_t)C * P.xstride * rs, &P.x1));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 3, (size_t)C * (P.bpc + 1), &P.pref));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 4, (size_t)C * P.bpc, &P.bsum));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 10, (size_t)C * P.bpc * 4, &P.blk_m));
  P.blk_s = P.blk_m + (size_t)C * P.bpc; P.blk_q = P.blk_s + (size_t)C * P.bpc; P.blk_x = P.blk_q + (size_t)C * P.bpc;
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 5, (size_t)C, &P.counter));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 6, (size_t)2 * C, &P.res));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 7, (size_t)2 * C, &P.seg));
  bool peer = false;
  if (sh) {
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 8, (size_t)C, &P.rec_local));
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 9, (size_t)C * sh->world, &P.rec_all));
    // the ranks' records travel through peer memory from k_st_step's tail when the group attached its inboxes
    // (bssm_shard_peer_attach); otherwise ncclAllGather + k_st_merge between the two kernels of an observation
    peer = ctx->peer_on && sh->world > 1 && sh->world <= BSSM_PEER_MAX_WORLD && C == 1;
    if (peer) {
      for (int g = 0; g < sh->world; g++) P.peer[g] = (StPeerSlot*)ctx->peer_ptr[g];
      const char* to_env = getenv("BSSM_PEER_TIMEOUT_MS");
      const long long to_ms = to_env ? atoll(to_env) : 0;
      P.peer_timeout_ns = (unsigned long long)(to_ms > 0 ? to_ms : 30000) * 1000000ull;
      P.peer_seq0 = ctx->peer_seq;
      ctx->peer_seq += (unsigned long long)L.T + 2;
    }
  }
  if (mn) {
    if (sh) { set_error("streaming engine: multinomial resampling is not available for the particle-sharded filter"); return BSSM_ERR_UNSUPPORTED; }
    // sorted uniforms from exponential spacings (bssm_stream.cuh): positions of the output slots + the scan of the spacings
    P.mn_nt = (cap + 1 + MN_TILE - 1) / MN_TILE + 1;
    // The positions of an observation's output slots depend on nothing but its index: they are laid out on a second stream
    // while the main stream still propagates (double-buffered by the observation's parity, for every observation, since its
    // resampling decision is not known yet) -- at N = 2^20 the three launches were half of an observation's time.
    // Only where launches, not bytes, are what an observation costs (C2: 22.1 -> 29.8 G particle-timesteps/s; at N = 2^24 the
    // positions of the 37 % of observations that do not resample are wasted bandwidth: 47.2 -> 42.9, profiles/r2_ab_multinomial.txt).
    // BSSM_ST_MN_AHEAD=0 / 1: in line on the main stream and only for the filters that resample / ahead.
    const char* e = getenv("BSSM_ST_MN_AHEAD");
    P.mn_ahead = e ? (atoi(e) != 0) : ((long long)C * cap <= (1LL << 22));
    const size_t rows = P.mn_ahead ? (size_t)2 * C : (size_t)C;
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 2, rows * P.xstride, &P.mn_pos));
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 8, rows * P.mn_nt, &P.mn_tsum));
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 9, rows, &P.mn_total));
    if (P.mn_ahead && !ctx->aux_stream) {
      BSSM_CK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
      BSSM_CK(cudaEventCreateWithFlags(&ctx->ev_mn_start, cudaEventDisableTiming));
      for (int i = 0; i < 2; i++) { BSSM_CK(cudaEventCreateWithFlags(&ctx->ev_mn_ready[i], cudaEventDisableTiming)); BSSM_CK(cudaEventCreateWithFlags(&ctx->ev_mn_free[i], cudaEventDisableTiming)); }
    }
  }
  P.dbg = nullptr;
#ifdef BSSM_ST_CHAIN_TIMING
  const bool dbg_ok = true;
#else
  const bool dbg_ok = !chain;
#endif
  if (getenv("BSSM_ST_TIMING") && dbg_ok) {
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 11, (size_t)8, &P.dbg));
    BSSM_CK(cudaMemsetAsync(P.dbg, 0, 8 * sizeof(long long), st));
  }
  if ((long long)P.bpc * C > 2147483647LL) { set_error("streaming engine: too many blocks (%d filters x %d)", C, P.bpc); return BSSM_ERR_UNSUPPORTED; }
  k_st_setup<<<(C + 127) / 128, 128, 0, st>>>(P, goff0, nloc0);
  BSSM_LAUNCH(ctx, "k_st_setup");
  const dim3 grid((unsigned int)((size_t)P.bpc * C));   // block index within the filter fastest
  {
    void* args[] = {&P};
    BSSM_CK(cudaLaunchKernel(K.init, grid, dim3(ST_THREADS), args, 0, st));
    BSSM_LAUNCH(ctx, "k_st_init");
  }
  const bool may_resample = f.ralg != BSSM_SIS;
  // the two kernels of an observation chain by programmatic dependent launch -- on one GPU, and for the sharded filter when its
  // records travel through peer memory from k_st_step's tail (with the NCCL exchange there is an all-gather and a merge kernel
  // between them: ordinary launches)
  const char* pdl_env = getenv("BSSM_ST_PDL");
  const bool pdl = (!sh || peer) && !(pdl_env && atoi(pdl_env) == 0);
  if (chain) {
    const int per = chain_slots / P.bpc;     // filters per cooperative launch
    for (int c0 = 0; c0 < C; c0 += per) {
      int cb = C - c0 < per ? C - c0 : per, T = L.T;
      void* args[] = {&P, &c0, &T};
      const cudaError_t ce = cudaLaunchCooperativeKernel(K.chain, dim3((unsigned int)((size_t)cb * P.bpc)), dim3(ST_THREADS), args, 0, st);
      if (ce == cudaErrorCooperativeLaunchTooLarge && c0 == 0) {   // fewer resident blocks than the occupancy query promised (a shared
        cudaGetLastError();                                         // GPU): nothing has run yet -- the launch-per-body form serves any bpc
        chain = false;
        break;
      }
      BSSM_CK(ce);
      BSSM_LAUNCH(ctx, "k_st_chain");
    }
  }
  auto mn_positions = [&](cudaStream_t s, int obs) -> int {
    const dim3 gmn((unsigned int)C, (unsigned int)P.mn_nt);
    k_st_mn_sums<<<gmn, MN_THREADS, 0, s>>>(P, obs);
    BSSM_LAUNCH(ctx, "k_st_mn_sums");
    k_st_mn_scan<<<C, MN_THREADS, 0, s>>>(P, obs);
    BSSM_LAUNCH(ctx, "k_st_mn_scan");
    k_st_mn_positions<<<gmn, MN_THREADS, 0, s>>>(P, obs);
    BSSM_LAUNCH(ctx, "k_st_mn_positions");
    return BSSM_OK;
  };
  if (mn && P.mn_ahead && may_resample && L.T > 0) {
    // the second stream starts behind everything already queued on the main one (an earlier filter's resampling may still read
    // the position buffers), with the positions of observation 0
    BSSM_CK(cudaEventRecord(ctx->ev_mn_start, st));
    BSSM_CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_mn_start, 0));
    BSSM_TRY(mn_positions(ctx->aux_stream, 0));
    BSSM_CK(cudaEventRecord(ctx->ev_mn_ready[0], ctx->aux_stream));
  }
  for (int obs = 0; obs < L.T && !chain; obs++) {
    BSSM_TRY(st_launch(ctx, K.step, grid, ST_THREADS, P, &obs, "k_st_step", pdl && obs > 0));
    if (sh && !peer) {
      BSSM_TRY(shard_allgather(ctx, sh, P.rec_local, P.rec_all, (size_t)C * sizeof(StRec)));
      k_st_merge<<<(C + 127) / 128, 128, 0, st>>>(P, obs);
      BSSM_LAUNCH(ctx, "k_st_merge");
    }
    if (may_resample) {
      if (mn && !P.mn_ahead) {   // the positions of this observation's output slots (the kernels return at once for filters that do not resample)
        BSSM_TRY(mn_positions(st, obs));
      } else if (mn) {
        // observation obs + 1's positions go out now (their buffer is free once the resampling of obs - 1 has run); this
        // observation's were laid out an observation ago
        if (obs + 1 < L.T) {
          if (obs >= 1) BSSM_CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_mn_free[(obs + 1) & 1], 0));
          BSSM_TRY(mn_positions(ctx->aux_stream, obs + 1));
          BSSM_CK(cudaEventRecord(ctx->ev_mn_ready[(obs + 1) & 1], ctx->aux_stream));
        }
        BSSM_CK(cudaStreamWaitEvent(st, ctx->ev_mn_ready[obs & 1], 0));
      }
      BSSM_TRY(st_launch(ctx, K.resample, grid, ST_THREADS, P, &obs, "k_st_resample", pdl && !mn));
      if (mn && P.mn_ahead) BSSM_CK(cudaEventRecord(ctx->ev_mn_free[obs & 1], st));
    }
  }
  if (P.dbg && chain) {
    long long h[8];
    BSSM_CK(cudaMemcpyAsync(h, P.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    BSSM_CK(cudaStreamSynchronize(st));
    const double nbk = h[3] > 0 ? (double)h[3] : 1.0;
    fprintf(stderr, "[bssm chain timing] per block, cycles: step body %.0f (of which waiting for the merge %.0f), resample body %.0f, wait after resampling %.0f; %lld blocks, bpc=%d\n",
            h[0] / nbk, h[4] / nbk, h[1] / nbk, h[2] / nbk, h[3], P.bpc);
  } else if (P.dbg) {
    long long h[8];
    BSSM_CK(cudaMemcpyAsync(h, P.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    BSSM_CK(cudaStreamSynchronize(st));
    fprintf(stderr, "[bssm stream timing] last k_st_step, merging block: %lld cycles until its ticket, merge %lld (max pass %lld, block max %lld, sum pass %lld, prefix pass + rest %lld), global bookkeeping %lld; bpc=%d nt=%d\n",
            h[0], h[1], h[4], h[5], h[6], h[7], h[2], P.bpc, P.nt);
  }
  k_st_flush<TS><<<C, 256, 0, st>>>(P, L.T);
  BSSM_LAUNCH(ctx, "k_st_flush");
  if (sh && !peer) {
    BSSM_TRY(shard_allgather(ctx, sh, P.rec_local, P.rec_all, (size_t)C * sizeof(StRec)));
    k_st_flush_merge<<<(C + 127) / 128, 128, 0, st>>>(P, L.T);
    BSSM_LAUNCH(ctx, "k_st_flush_merge");
  }
  return BSSM_OK;
}

template <typename Model, typename Real, int PPT, int THREADS> static StreamKernels builtin_stream_kernels() {
  StreamKernels K;
  K.init = (void*)k_st_init<Model, Real, PPT, THREADS>; K.step = (void*)k_st_step<Model, Real, PPT, THREADS>;
  K.resample = (void*)k_st_resample<Model, Real, PPT, THREADS>;
  K.chain = (void*)k_st_chain<Model, Real, PPT, THREADS>;
  return K;
}
// 128 threads per block for batches, 256 for a few big filters (see the note in bssm_stream.cuh)
static bool stream_small_blocks(const FilterDev& f) {
  if (const char* e = getenv("BSSM_ST_THREADS")) return atoi(e) == 128;
  return f.C >= 16;
}
template <typename Model>
static int stream_model(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, const ShardRun* sh) {
  const bool small = stream_small_blocks(f);
  if (L.precision == BSSM_F64)
    return small ? stream_launch<double, 4, 128>(ctx, f, L, sh, builtin_stream_kernels<Model, double, 4, 128>())
                 : stream_launch<double, 4, 256>(ctx, f, L, sh, builtin_stream_kernels<Model, double, 4, 256>());
  return small ? stream_launch<float, 8, 128>(ctx, f, L, sh, builtin_stream_kernels<Model, float, 8, 128>())
               : stream_launch<float, 8, 256>(ctx, f, L, sh, builtin_stream_kernels<Model, float, 8, 256>());
}

int stream_filter_enqueue(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, const ShardRun* sh) {
  if (L.model >= BSSM_USER_MODEL_BASE) {
    const UserModelInfo* u = user_model(ctx, L.model);
    if (!u || !u->stream_ok) { set_error("streaming engine: user model %d not supported (1-D state, one normal per init / transition)", L.model); return BSSM_ERR_UNSUPPORTED; }
    const int v = stream_small_blocks(f) ? 1 : 0;
    if (L.precision == BSSM_F64) return v ? stream_launch<double, 4, 128>(ctx, f, L, sh, u->s64[1]) : stream_launch<double, 4, 256>(ctx, f, L, sh, u->s64[0]);
    return v ? stream_launch<float, 8, 128>(ctx, f, L, sh, u->s32[1]) : stream_launch<float, 8, 256>(ctx, f, L, sh, u->s32[0]);
  }
  switch (L.model) {
    case BSSM_MODEL_AR_SIN: return stream_model<ModelArSin>(ctx, f, L, sh);
    case BSSM_MODEL_LG: return stream_model<ModelLG>(ctx, f, L, sh);
    case BSSM_MODEL_AR_COS: return stream_model<ModelArCos>(ctx, f, L, sh);
    case BSSM_MODEL_RW_DRIFT: return stream_model<ModelRwDrift>(ctx, f, L, sh);
  }
  set_error("streaming engine: model %d not supported", L.model);
  return BSSM_ERR_UNSUPPORTED;
}

}  // namespace bssm
