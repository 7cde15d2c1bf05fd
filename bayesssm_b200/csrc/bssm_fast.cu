// bssm_fast.cu -- host side of the persistent bootstrap-filter kernel (bssm_fast.cuh).
#include "bssm_engine.cuh"
#include "bssm_fast.cuh"

#include <stdlib.h>

namespace bssm {

bool fast_supported(const FilterDev& f, const FilterLaunch& L) {
  if (f.algorithm != BSSM_BPF || L.hist || f.noise.injected || f.anc_history || f.carry) return false;
  if (L.resample_fn == BSSM_MULTINOMIAL) return false;
  if (!(L.model == BSSM_MODEL_AR_SIN || L.model == BSSM_MODEL_LG || L.model == BSSM_MODEL_AR_COS || L.model == BSSM_MODEL_RW_DRIFT)) return false;
  if ((long long)f.N > (long long)FAST_MAX_G * FAST_MAX_NB) return false;
  return true;
}

static int fast_ppt_choice(int nb_max) {
  const char* e = getenv("BSSM_FAST_PPT");
  if (e && atoi(e) == 8) return 8;
  if (e && atoi(e) == 16) return 16;
  return nb_max >= 2048 ? 16 : 8;
}

template <typename Model, typename Real, int PPT, bool HEADS>
static int fast_launch(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L, int G) {
  const int nsm = ctx->prop.multiProcessorCount;
  int nb_max = (f.N + G - 1) / G;
  nb_max = (nb_max + PPT - 1) / PPT * PPT;
  if (nb_max > FAST_MAX_NB) { set_error("persistent kernel: %d particles per CTA exceed %d", nb_max, FAST_MAX_NB); return BSSM_ERR_UNSUPPORTED; }
  int threads = (nb_max / PPT + 31) / 32 * 32;
  if (threads < 32) threads = 32;
  // staging capacity of one expansion pass: PPT * 5/4 output slots per thread (25 % beyond the slice; more offspring
  // than that take further passes)
  const int cap = HEADS ? threads * fast_spt(PPT) : (nb_max + FAST_SLACK + 31) / 32 * 32;
  size_t smem = (size_t)((5 * G + 1) & ~1) * sizeof(double) + 5 * 32 * sizeof(double) + (size_t)cap * sizeof(Real) + (size_t)cap * sizeof(unsigned int);
  if (HEADS) smem += (size_t)cap * sizeof(unsigned int) + (size_t)threads * PPT * sizeof(Real);   // head array + the CTA's particles
  auto kern = k_fast_bpf<Model, Real, PPT, HEADS>;
  BSSM_CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  BSSM_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  long long resident = (long long)per_sm * nsm;
  if (resident < G) { set_error("persistent kernel: group of %d CTAs does not fit (%lld resident)", G, resident); return BSSM_ERR_UNSUPPORTED; }
  if (getenv("BSSM_TEST_FAST_LAUNCH_UNSUPPORTED")) { set_error("persistent kernel: launch refused (BSSM_TEST_FAST_LAUNCH_UNSUPPORTED: test hook of the AUTO fallback)"); return BSSM_ERR_UNSUPPORTED; }
  int ngroups = (int)(resident / G);
  if (ngroups > f.C) ngroups = f.C;
  FastParams P;
  P.f = f; P.G = G; P.ngroups = ngroups; P.resample_fn = L.resample_fn; P.nb_max = nb_max; P.cap = cap;
  BSSM_TRY(scratch(ctx, SL_FAST_BASE + 0, (size_t)ngroups * 2 * G, &P.rec));
  // x_new holds LL elements (value + epoch tag): 8 bytes (f32) / 16 bytes (f64) per particle; tags start at 0
  const size_t xbytes = (size_t)ngroups * G * nb_max * (sizeof(Real) == 4 ? 8 : 16);
  BSSM_TRY(scratch_get(ctx, SL_FAST_BASE + 2, xbytes, &P.xnew));
  BSSM_CK(cudaMemsetAsync(P.rec, 0, sizeof(FastRec) * (size_t)ngroups * 2 * G, ctx->stream));
  BSSM_CK(cudaMemsetAsync(P.xnew, 0, xbytes, ctx->stream));
  P.timing = nullptr;
  const bool timing = getenv("BSSM_FAST_TIMING") != nullptr;
  if (timing) {
    BSSM_TRY(scratch(ctx, SL_FAST_BASE + 3, (size_t)ngroups * G * 16, &P.timing));
    BSSM_CK(cudaMemsetAsync(P.timing, 0, sizeof(long long) * (size_t)ngroups * G * 16, ctx->stream));
  }
  void* args[] = {&P};
  BSSM_CK(cudaLaunchCooperativeKernel((void*)kern, dim3(ngroups * G), dim3(threads), args, smem, ctx->stream));
  BSSM_LAUNCH(ctx, "k_fast_bpf");
  if (timing) {   // diagnostics only: per-phase cycles of thread 0, averaged over the CTAs
    std::vector<long long> h((size_t)ngroups * G * 16);
    BSSM_CK(cudaMemcpyAsync(h.data(), P.timing, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    BSSM_CK(cudaStreamSynchronize(ctx->stream));
    const char* names[9] = {"loop head", "P1 propagate/weights/reduce", "next-step normals", "B1 poll", "P2", "stage uniforms",
                            "offspring ranges", "scatter+copy-out", "B2 reload"};
    double tot = 0;
    double avg[9];
    for (int i = 0; i < 9; i++) { double a = 0; for (int c = 0; c < ngroups * G; c++) a += (double)h[(size_t)c * 16 + i]; avg[i] = a / (ngroups * G); tot += avg[i]; }
    fprintf(stderr, "[bssm fast timing] G=%d groups=%d threads=%d T=%d: cycles per observation (thread 0, mean over CTAs)\n", G, ngroups, threads, f.T);
    for (int i = 0; i < 9; i++) fprintf(stderr, "  %-28s %9.0f  (%4.1f%%)\n", names[i], avg[i] / (f.T > 0 ? f.T : 1), 100.0 * avg[i] / tot);
  }
  return BSSM_OK;
}

template <typename Model>
static int fast_model(bssm_ctx* ctx, FilterDev& f, const FilterLaunch& L) {
  const int nsm = ctx->prop.multiProcessorCount;
  // group size: as few CTAs as hold the particles, widened to fill the chip when there are few filters
  // many filters: half-size slices put two groups on every SM, so one group's waits overlap the other's work
  // (measured +9 % on 128 chains x 65536); a single big filter keeps one CTA per SM
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
  // expansion of the offspring: head flags + running maximum on small slices (8 particles per thread), per-source
  // scatter loops on big ones (profiles/r1_ab_experiments.md)
  if (L.precision == BSSM_F64) return nb <= 2048 ? fast_launch<Model, double, 8, true>(ctx, f, L, G) : fast_launch<Model, double, 8, false>(ctx, f, L, G);
  if (fast_ppt_choice(nb) == 16) return fast_launch<Model, float, 16, false>(ctx, f, L, G);
  return fast_launch<Model, float, 8, true>(ctx, f, L, G);
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
