// bssm_stream.cu -- host side of the streaming bootstrap-filter engine (bssm_stream.cuh): two kernel
// launches per observation, all decisions on the device, no host synchronisation inside a filter.
// With a shard communicator (bssm_shard.cu) one ncclAllGather of a 64-byte record per filter and a
// one-thread merge kernel sit between the two kernels of an observation.
#include "bssm_engine.cuh"
#include "bssm_stream.cuh"

#include <math.h>
#include <stdlib.h>

namespace bssm {

bool stream_supported(const FilterDev& f, const FilterLaunch& L) {
  if (f.algorithm != BSSM_BPF || L.hist || f.noise.injected || f.anc_history) return false;
  if (L.resample_fn == BSSM_MULTINOMIAL) return false;
  if (!(L.model == BSSM_MODEL_AR_SIN || L.model == BSSM_MODEL_LG || L.model == BSSM_MODEL_AR_COS || L.model == BSSM_MODEL_RW_DRIFT)) return false;
  return true;
}

template <typename Model, typename Real, int PPT>
static int stream_launch(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, const ShardRun* sh) {
  constexpr int TS = ST_THREADS * PPT;
  cudaStream_t st = ctx->stream;
  StreamParams P;
  memset(&P, 0, sizeof(P));
  P.f = f; P.resample_fn = L.resample_fn;
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
  BSSM_TRY(scratch_get(ctx, SL_ST_BASE + 0, (size_t)C * P.xstride * rs, &P.x0));
  BSSM_TRY(scratch_get(ctx, SL_ST_BASE + 1, (size_t)C * P.xstride * rs, &P.x1));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 2, (size_t)C * P.nt * 4, &P.part_m));
  P.part_s = P.part_m + (size_t)C * P.nt; P.part_q = P.part_s + (size_t)C * P.nt; P.part_x = P.part_q + (size_t)C * P.nt;
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 3, (size_t)C * (P.nt + 1), &P.pref));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 4, (size_t)C * P.nt, &P.bsum));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 5, (size_t)C, &P.counter));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 6, (size_t)2 * C, &P.res));
  BSSM_TRY(scratch(ctx, SL_ST_BASE + 7, (size_t)2 * C, &P.seg));
  if (sh) {
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 8, (size_t)C, &P.rec_local));
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 9, (size_t)C * sh->world, &P.rec_all));
  }
  P.dbg = nullptr;
  if (getenv("BSSM_ST_TIMING")) {
    BSSM_TRY(scratch(ctx, SL_ST_BASE + 11, (size_t)8, &P.dbg));
    BSSM_CK(cudaMemsetAsync(P.dbg, 0, 8 * sizeof(long long), st));
  }
  if ((long long)P.nt * C > 2147483647LL) { set_error("streaming engine: too many tiles (%d filters x %d)", C, P.nt); return BSSM_ERR_UNSUPPORTED; }
  k_st_setup<<<(C + 127) / 128, 128, 0, st>>>(P, goff0, nloc0);
  BSSM_LAUNCH(ctx, "k_st_setup");
  // blocks per filter: the resident block slots of the chip split over the filters, each block walking
  // several tiles; among a few candidates take the one with the fewest rounds (waves x tiles per block)
  auto pick_bpc = [&](const void* kern, const char* env) -> int {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ST_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    long long slots = (long long)per_sm * ctx->prop.multiProcessorCount;
    if (const char* e = getenv(env)) { int v = atoi(e); if (v >= 1) return v > P.nt ? P.nt : v; }
    long long b0 = (slots + C - 1) / C;
    if (b0 > P.nt) b0 = P.nt;
    long long best = b0, best_cost = -1;
    for (long long b = b0; b <= P.nt && b <= 4 * b0 + 3; b++) {
      const long long waves = ((long long)C * b + slots - 1) / slots;
      const long long cost = waves * ((P.nt + b - 1) / b) * 16 + waves;   // + a little per wave for the block prologue
      if (best_cost < 0 || cost < best_cost) { best = b; best_cost = cost; }
    }
    return (int)best;
  };
  P.bpc = pick_bpc((const void*)k_st_step<Model, Real, PPT>, "BSSM_ST_BPC");
  P.bpc_r = pick_bpc((const void*)k_st_resample<Model, Real, PPT>, "BSSM_ST_BPC_R");
  const dim3 grid_init((unsigned int)((size_t)P.nt * C));   // tile index fastest
  const dim3 grid((unsigned int)((size_t)P.bpc * C)), grid_r((unsigned int)((size_t)P.bpc_r * C));
  k_st_init<Model, Real, PPT><<<grid_init, ST_THREADS, 0, st>>>(P);
  BSSM_LAUNCH(ctx, "k_st_init");
  const bool may_resample = f.ralg != BSSM_SIS;
  for (int obs = 0; obs < L.T; obs++) {
    k_st_step<Model, Real, PPT><<<grid, ST_THREADS, 0, st>>>(P, obs);
    BSSM_LAUNCH(ctx, "k_st_step");
    if (sh) {
      BSSM_TRY(shard_allgather(ctx, sh, P.rec_local, P.rec_all, (size_t)C * sizeof(StRec)));
      k_st_merge<<<(C + 127) / 128, 128, 0, st>>>(P, obs);
      BSSM_LAUNCH(ctx, "k_st_merge");
    }
    if (may_resample) {
      k_st_resample<Model, Real, PPT><<<grid_r, ST_THREADS, 0, st>>>(P, obs);
      BSSM_LAUNCH(ctx, "k_st_resample");
    }
  }
  if (P.dbg) {
    long long h[8];
    BSSM_CK(cudaMemcpyAsync(h, P.dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    BSSM_CK(cudaStreamSynchronize(st));
    fprintf(stderr, "[bssm stream timing] last k_st_step, merging block: %lld cycles until its ticket, merge %lld (max pass %lld, block max %lld, sum pass %lld, prefix pass + rest %lld), global bookkeeping %lld; bpc=%d bpc_r=%d nt=%d\n",
            h[0], h[1], h[4], h[5], h[6], h[7], h[2], P.bpc, P.bpc_r, P.nt);
  }
  k_st_flush<<<C, ST_THREADS, 0, st>>>(P, L.T, TS);
  BSSM_LAUNCH(ctx, "k_st_flush");
  if (sh) {
    BSSM_TRY(shard_allgather(ctx, sh, P.rec_local, P.rec_all, (size_t)C * sizeof(StRec)));
    k_st_flush_merge<<<(C + 127) / 128, 128, 0, st>>>(P, L.T);
    BSSM_LAUNCH(ctx, "k_st_flush_merge");
  }
  return BSSM_OK;
}

template <typename Model>
static int stream_model(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, const ShardRun* sh) {
  if (L.precision == BSSM_F64) return stream_launch<Model, double, 4>(ctx, f, L, sh);
  return stream_launch<Model, float, 8>(ctx, f, L, sh);
}

int stream_filter_enqueue(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, const ShardRun* sh) {
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
