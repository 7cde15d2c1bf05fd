// bssm_pmmh.cu -- placeholder until the device-resident PMMH lands (next commit).
#include "bssm_engine.cuh"
#include <math.h>
using namespace bssm;
extern "C" {
int bssm_pmmh_run(bssm_ctx*, const bssm_pmmh_config*, const double*, const double*, bssm_pmmh_result*) {
  set_error("bssm_pmmh_run: not built yet"); return BSSM_ERR_UNSUPPORTED;
}
int bssm_filter_run_device(bssm_ctx*, const bssm_filter_config*, const double*, const double*, double*, float*) {
  set_error("bssm_filter_run_device: not built yet"); return BSSM_ERR_UNSUPPORTED;
}
int bssm_model_compile(bssm_ctx*, const char*, int*) { set_error("bssm_model_compile: not built yet"); return BSSM_ERR_UNSUPPORTED; }
const char* bssm_model_compile_log(bssm_ctx* ctx) { return ctx ? ctx->compile_log.c_str() : ""; }
double bssm_transform(double th, int tr) { return tr == BSSM_TR_LOG ? log(th) : (tr == BSSM_TR_LOGIT ? log(th / (1.0 - th)) : th); }
double bssm_back_transform(double z, int tr) { return tr == BSSM_TR_LOG ? exp(z) : (tr == BSSM_TR_LOGIT ? 1.0 / (1.0 + exp(-z)) : z); }
double bssm_log_jacobian(const double* th, const int* tr, int p) {
  double s = 0; for (int j = 0; j < p; j++) { if (tr[j] == BSSM_TR_LOG) s += log(th[j]); else if (tr[j] == BSSM_TR_LOGIT) s += log(1.0 / (th[j] * (1.0 - th[j]))); } return s;
}
double bssm_log_prior(int kind, double a, double b, double x) {
  const double LSP = 0.918938533204672741780329736406;
  switch (kind) {
    case BSSM_PRIOR_FLAT: return 0.0;
    case BSSM_PRIOR_NORMAL: { double z = (x - a) / b; return -(LSP + 0.5 * z * z + log(b)); }
    case BSSM_PRIOR_EXP: return x < 0 ? -INFINITY : log(a) - a * x;
    case BSSM_PRIOR_UNIF: return (a <= x && x <= b) ? -log(b - a) : -INFINITY;
    case BSSM_PRIOR_HALFNORMAL: { if (x < 0) return -INFINITY; double z = x / a; return log(2.0) - (LSP + 0.5 * z * z + log(a)); }
  }
  return NAN;
}
}
