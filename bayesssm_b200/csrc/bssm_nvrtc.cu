// bssm_nvrtc.cu -- user models as CUDA device-function snippets compiled by NVRTC.
// Placeholder: the runtime-compilation path lands after the built-in models are parity-green.
#include "bssm_engine.cuh"
using namespace bssm;
extern "C" {
int bssm_model_compile(bssm_ctx*, const char*, int*) {
  set_error("bssm_model_compile: NVRTC user models are not available in this build");
  return BSSM_ERR_UNSUPPORTED;
}
const char* bssm_model_compile_log(bssm_ctx* ctx) { return ctx ? ctx->compile_log.c_str() : ""; }
}
