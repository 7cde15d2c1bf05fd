// bssm_diag.cu -- C ABI of the device MCMC diagnostics (kernels in bssm_diag.cuh).
#include "bssm_engine.cuh"
#include "bssm_diag.cuh"

using namespace bssm;

extern "C" int bssm_mcmc_diagnostics(bssm_ctx* ctx, const double* draws, int k, int m_total, int p, int burn_in,
                                     double* ess, double* rhat, int32_t* flags, float* device_ms) {
  if (!ctx || !draws || k < 1 || p < 1 || m_total < 1 || burn_in < 0 || (!ess && !rhat)) {
    set_error("bssm_mcmc_diagnostics: bad argument");
    return BSSM_ERR_BAD_ARG;
  }
  const int m = m_total - burn_in;
  if (m < 2) { set_error("Number of iterations must be at least 2."); return BSSM_ERR_BAD_ARG; }         // R/ESS.R:35-37, R/rhat.R:30-32
  if (ess && k < 2) { set_error("Number of chains must be at least 2."); return BSSM_ERR_BAD_ARG; }      // R/ESS.R:38-40
  BSSM_CK(cudaSetDevice(ctx->device));
  const size_t n_in = (size_t)k * m_total * p, n_ser = (size_t)p * k * m;
  double *dx, *xc, *mom, *acov, *small;
  int* dflags;
  BSSM_TRY(scratch(ctx, SL_DG_BASE + 0, n_in, &dx));
  BSSM_TRY(scratch(ctx, SL_DG_BASE + 1, n_ser, &xc));
  BSSM_TRY(scratch(ctx, SL_DG_BASE + 2, (size_t)p * k * 6, &mom));
  BSSM_TRY(scratch(ctx, SL_DG_BASE + 3, ess ? n_ser : 1, &acov));
  BSSM_TRY(scratch(ctx, SL_DG_BASE + 4, (size_t)p * 4 + (size_t)p * m, &small));   // par[p][2], ess[p], rhat[p], rho[p][m]
  BSSM_TRY(scratch(ctx, SL_DG_BASE + 5, (size_t)p, &dflags));
  BSSM_CK(cudaMemcpyAsync(dx, draws, sizeof(double) * n_in, cudaMemcpyHostToDevice, ctx->stream));
  DiagArgs a;
  a.x = dx + (size_t)burn_in * p;
  a.chain_stride = (long long)m_total * p;
  a.iter_stride = p;
  a.k = k; a.m = m; a.p = p;
  a.xc = xc; a.mom = mom; a.acov = acov;
  a.par = small;
  a.ess = ess ? small + (size_t)p * 2 : nullptr;
  a.rhat = small + (size_t)p * 3;
  a.rho = small + (size_t)p * 4;
  a.flags = dflags;
  auto blocks = [](long long n) { return (unsigned)((n + 255) / 256); };
  BSSM_CK(cudaEventRecord(ctx->ev0, ctx->stream));
  k_diag_moments<<<blocks(diag_n_moments(a)), 256, 0, ctx->stream>>>(a);
  BSSM_LAUNCH(ctx, "k_diag_moments");
  if (ess) {
    k_diag_acov<<<blocks(diag_n_acov(a)), 256, 0, ctx->stream>>>(a);
    BSSM_LAUNCH(ctx, "k_diag_acov");
  }
  k_diag_between<<<blocks(p), 256, 0, ctx->stream>>>(a);
  BSSM_LAUNCH(ctx, "k_diag_between");
  if (ess) {
    k_diag_rho<<<blocks(diag_n_rho(a)), 256, 0, ctx->stream>>>(a);
    BSSM_LAUNCH(ctx, "k_diag_rho");
    k_diag_geyer<<<blocks(p), 256, 0, ctx->stream>>>(a);
    BSSM_LAUNCH(ctx, "k_diag_geyer");
  }
  BSSM_CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (ess) BSSM_CK(cudaMemcpyAsync(ess, a.ess, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
  if (rhat) BSSM_CK(cudaMemcpyAsync(rhat, a.rhat, sizeof(double) * p, cudaMemcpyDeviceToHost, ctx->stream));
  if (flags) BSSM_CK(cudaMemcpyAsync(flags, dflags, sizeof(int32_t) * p, cudaMemcpyDeviceToHost, ctx->stream));
  BSSM_CK(cudaStreamSynchronize(ctx->stream));
  if (device_ms) BSSM_CK(cudaEventElapsedTime(device_ms, ctx->ev0, ctx->ev1));
  return BSSM_OK;
}
