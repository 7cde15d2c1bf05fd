// bssm_fast.cu -- host side of the persistent bootstrap-filter kernel (bssm_fast.cuh).
#include "bssm_engine.cuh"
#include "bssm_fast.cuh"

#include <stdlib.h>

namespace bssm {

bool fast_supported(const FilterDev& f, const FilterLaunch& L) {
  if (f.algorithm != BSSM_BPF || L.hist || f.noise.injected || f.anc_history) return false;
  if (L.resample_fn == BSSM_MULTINOMIAL) return false;
  if (!(L.model == BSSM_MODEL_AR_SIN || L.model == BSSM_MODEL_LG || L.model == BSSM_MODEL_AR_COS || L.model == BSSM_MODEL_RW_DRIFT)) return false;
  if ((long long)f.N > (long long)FAST_MAX_G * FAST_MAX_NB) return false;
  return true;
}

static int fast_ppt_choice(int nb_max) {
  const char* e = getenv("BSSM_FAST_PPT");
  if (e && atoi(e) == 8) return 8;
  if (e && atoi(e) == 16) return 16;
  if (e && atoi(e) == 12 && nb_max >= 2048) return 12;
  return nb_max >= 2048 ? 16 : 8;
}

template <typename Model, typename Real, int PPT, int NWMAX>
static int fast_launch(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, int G) {
  const int nsm = ctx->prop.multiProcessorCount;
  int uw_req = -1;
  if (const char* e = getenv("BSSM_FAST_UW")) uw_req = atoi(e);
  const FastGeom g = fast_geometry<Real, PPT>(f.N, G, uw_req);
  if (g.nb_max > FAST_MAX_NB || g.nw > NWMAX) { set_error("persistent kernel: %d particles per CTA exceed %d", g.nb_max, NWMAX * 32 * PPT); return BSSM_ERR_UNSUPPORTED; }
  auto kern = k_fast_bpf<Model, Real, PPT, NWMAX>;
  BSSM_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  int per_sm = 0;
  BSSM_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, g.threads, g.smem));
  long long resident = (long long)per_sm * nsm;
  if (resident < G) { set_error("persistent kernel: group of %d CTAs does not fit (%lld resident)", G, resident); return BSSM_ERR_UNSUPPORTED; }
  int ngroups = (int)(resident / G);
  if (ngroups > f.C) ngroups = f.C;
  FastParams P;
  P.f = f; P.G = G; P.ngroups = ngroups; P.resample_fn = L.resample_fn; P.nb_max = g.nb_max; P.xstride = g.xstride; P.ucap = g.ucap; P.uw = g.uw;
  const size_t rec_units = fast_rec_units(ngroups, G, FastRecLayout<sizeof(Real) == 4>::NUS);
  BSSM_TRY(scratch(ctx, SL_FAST_BASE + 0, rec_units, &P.rec));
  // x_new holds LL elements (value + epoch tag): 8 bytes (f32) / 16 bytes (f64) per particle; tags start at 0
  const size_t xbytes = (size_t)ngroups * g.xstride * (sizeof(Real) == 4 ? 8 : 16);
  BSSM_TRY(scratch_get(ctx, SL_FAST_BASE + 2, xbytes, &P.xnew));
  BSSM_CK(cudaMemsetAsync(P.rec, 0, sizeof(uint4) * rec_units, ctx->stream));
  BSSM_CK(cudaMemsetAsync(P.xnew, 0, xbytes, ctx->stream));
  P.timing = nullptr;
  const bool timing = getenv("BSSM_FAST_TIMING") != nullptr;
  if (timing) {
    BSSM_TRY(scratch(ctx, SL_FAST_BASE + 3, (size_t)ngroups * G * 192, &P.timing));
    BSSM_CK(cudaMemsetAsync(P.timing, 0, sizeof(long long) * (size_t)ngroups * G * 192, ctx->stream));
  }
  void* args[] = {&P};
  BSSM_CK(cudaLaunchCooperativeKernel((void*)kern, dim3(ngroups * G), dim3(g.threads), args, g.smem, ctx->stream));
  BSSM_LAUNCH(ctx, "k_fast_bpf");
  if (timing) {   // diagnostics only (build with -DBSSM_FAST_TIMING_BUILD): per-phase cycles of every warp, averaged over the CTAs
    std::vector<long long> h((size_t)ngroups * G * 192);
    BSSM_CK(cudaMemcpyAsync(h.data(), P.timing, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    BSSM_CK(cudaStreamSynchronize(ctx->stream));
    const int ncta = ngroups * G, T = f.T > 0 ? f.T : 1;
    fprintf(stderr, "[bssm fast timing] G=%d groups=%d threads=%d T=%d: cycles per observation, mean over CTAs\n"
                    "  0 reload | 11 P1a | 8 barriers(max,totals) | 1 P1b | 9 scan+publish | 2 overlap | 3 poll | 10 barrier(records) | 4 merge | 5 uniforms | 6 ranges | 7 expansion\n", G, ngroups, g.threads, f.T);
    for (int w = 0; w < g.nw && w < 16; w++) {
      fprintf(stderr, "  warp %2d", w);
      const int order[12] = {0, 11, 8, 1, 9, 2, 3, 10, 4, 5, 6, 7};
      for (int k = 0; k < 12; k++) { const int i = order[k]; double a = 0; for (int c = 0; c < ncta; c++) a += (double)h[((size_t)c * 16 + w) * 12 + i]; fprintf(stderr, " %7.0f", a / ncta / T); }
      fprintf(stderr, "\n");
    }
  }
  return BSSM_OK;
}

template <typename Model>
static int fast_model(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L) {
  const int nsm = ctx->prop.multiProcessorCount;
  // group size: as few CTAs as hold the particles, widened to fill the chip when there are few filters
  // many filters: half-size slices put two groups on every SM, so one group's waits overlap the other's work;
  // a single big filter keeps one CTA per SM
  int nb_cap = ((long long)f.C * ((f.N + FAST_MAX_NB - 1) / FAST_MAX_NB) >= nsm) ? FAST_MAX_NB / 2 : FAST_MAX_NB;
  if (const char* e = getenv("BSSM_FAST_NB")) { int v = atoi(e); if (v >= 256 && v <= FAST_MAX_NB) nb_cap = v; }
  int G = (f.N + nb_cap - 1) / nb_cap;
  if (f.C * G < nsm) {
    int wide = nsm / f.C;
    int cap = (f.N + 1023) / 1024;   // keep >= ~1024 particles per CTA
    if (wide > cap) wide = cap;
    if (wide > G) G = wide;
  }
  if (G > FAST_MAX_G) G = FAST_MAX_G;
  if (G < 1) G = 1;
  const int nb = (f.N + G - 1) / G;
  if (L.precision == BSSM_F64) return fast_launch<Model, double, 8, 28>(ctx, f, L, G);
  if (fast_ppt_choice(nb) == 16) return fast_launch<Model, float, 16, 14>(ctx, f, L, G);
  if (fast_ppt_choice(nb) == 12) return fast_launch<Model, float, 12, 19>(ctx, f, L, G);
  return nb <= 2048 ? fast_launch<Model, float, 8, 8>(ctx, f, L, G) : fast_launch<Model, float, 8, 28>(ctx, f, L, G);
}

int fast_filter_enqueue(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L) {
  switch (L.model) {
    case BSSM_MODEL_AR_SIN: return fast_model<ModelArSin>(ctx, f, L);
    case BSSM_MODEL_LG: return fast_model<ModelLG>(ctx, f, L);
    case BSSM_MODEL_AR_COS: return fast_model<ModelArCos>(ctx, f, L);
    case BSSM_MODEL_RW_DRIFT: return fast_model<ModelRwDrift>(ctx, f, L);
  }
  set_error("persistent kernel: model %d not supported", L.model);
  return BSSM_ERR_UNSUPPORTED;
}

}  // namespace bssm
