"""Compile a user-model snippet with NVRTC exactly as bssm_model_compile does -- the general engine's model kernels
and the streaming engine's kernels, for sm_100a -- without a GPU: catches constructs NVRTC rejects (host headers in
the embedded text, for instance) before any GPU time is spent.  tests/test_abi.py runs it."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bayesssm_b200", "csrc")
USER = open(sys.argv[1]).read() if len(sys.argv) > 1 else r'''
// stochastic volatility: x_t = mu + phi (x_{t-1} - mu) + sigma v_t,  y_t ~ N(0, exp(x_t))
struct UserModel {
  static constexpr int D = 1, NTHETA = 3, NCONST = 0, NZ_INIT = 1, NU_INIT = 0, NZ_TRANS = 1, NU_TRANS = 0,
                       NZ_MOVE = 0, NU_MOVE = 0, NPAR = 4;
  static constexpr bool HAS_AUX = false, HAS_MOVE = false;
  template <typename R> static BSSM_DEV void prepare(const double* th, R* par) {
    par[0] = (R)th[0]; par[1] = (R)th[1]; par[2] = (R)th[2]; par[3] = (R)(th[2] / sqrt(1.0 - th[1] * th[1]));
  }
  template <typename R> static BSSM_DEV void init(R* x, const R* par, const R* z, const double*) { x[0] = par[0] + par[3] * z[0]; }
  template <typename R> static BSSM_DEV void transition(R* x, const R* par, int, const R* z, const double*) {
    x[0] = par[0] + par[1] * (x[0] - par[0]) + par[2] * z[0];
  }
  template <typename R> static BSSM_DEV R loglik(const double* y, const R* x, const R*, int) {
    R yy = (R)y[0];
    return -((R)0.918938533204672741780329736406 + (R)0.5 * x[0] + (R)0.5 * yy * yy * Math<R>::exp_(-x[0]));
  }
  template <typename R> static BSSM_DEV R aux_loglik(const double* y, const R* x, const R* par, int t) { return loglik<R>(y, x, par, t); }
  template <typename R> static BSSM_DEV void move(R*, const double*, const R*, int, const R*, const double*) {}
};
'''
prog = '#include "bssm_filter.cuh"\n#include "bssm_stream.cuh"\nnamespace bssm {\n#line 1 "user_model.cu"\n' + USER + "\n}\n" + '''
namespace bssm {
template __global__ void k_st_init<UserModel, float, 8, 256>(StreamParams);
template __global__ void k_st_init<UserModel, double, 4, 256>(StreamParams);
template __global__ void k_st_step<UserModel, float, 8, 256>(const __grid_constant__ StreamParams, int);
template __global__ void k_st_step<UserModel, double, 4, 256>(const __grid_constant__ StreamParams, int);
template __global__ void k_st_resample<UserModel, float, 8, 256>(const __grid_constant__ StreamParams, int);
template __global__ void k_st_resample<UserModel, double, 4, 256>(const __grid_constant__ StreamParams, int);
template __global__ void k_st_init<UserModel, float, 8, 128>(StreamParams);
template __global__ void k_st_init<UserModel, double, 4, 128>(StreamParams);
template __global__ void k_st_step<UserModel, float, 8, 128>(const __grid_constant__ StreamParams, int);
template __global__ void k_st_step<UserModel, double, 4, 128>(const __grid_constant__ StreamParams, int);
template __global__ void k_st_resample<UserModel, float, 8, 128>(const __grid_constant__ StreamParams, int);
template __global__ void k_st_resample<UserModel, double, 4, 128>(const __grid_constant__ StreamParams, int);
template __global__ void k_init<UserModel, float>(FilterDev);
template __global__ void k_init<UserModel, double>(FilterDev);
template __global__ void k_weight<UserModel, float>(FilterDev, int, int, int);
template __global__ void k_weight<UserModel, double>(FilterDev, int, int, int);
template __global__ void k_post<UserModel, float>(FilterDev, int);
template __global__ void k_post<UserModel, double>(FilterDev, int);
}
extern "C" __global__ void bssm_user_dims(int* o) { o[0] = bssm::UserModel::D; }
'''
nv = C.CDLL("libnvrtc.so.12")
names = [b"bssm_common.cuh", b"bssm_models.cuh", b"bssm_filter.cuh", b"bssm_slots.cuh", b"bssm_stream.cuh"]
srcs = [open(os.path.join(CSRC, n.decode())).read().encode() for n in names]
p = C.c_void_p()
arr = (C.c_char_p * len(names))
assert nv.nvrtcCreateProgram(C.byref(p), prog.encode(), b"u.cu", len(names), arr(*srcs), arr(*names)) == 0
opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-lineinfo"]
rc = nv.nvrtcCompileProgram(p, len(opts), (C.c_char_p * len(opts))(*opts))
n = C.c_size_t()
nv.nvrtcGetProgramLogSize(p, C.byref(n))
buf = C.create_string_buffer(n.value)
nv.nvrtcGetProgramLog(p, buf)
print("rc", rc)
print(buf.value.decode()[:3000])
nv.nvrtcGetCUBINSize(p, C.byref(n))
print("cubin bytes", n.value)
sys.exit(0 if rc == 0 else 1)
