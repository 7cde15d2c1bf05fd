"""ctypes binding of libbayesssm_b200.so (the C ABI declared in include/bayesssm_b200.h).

The library is built in-tree by ``bayesssm_b200.build.build_native()`` (nvcc, sm_100a) and is the
only compute path of this package: if it is missing, or no CUDA device is present, every
entry point raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSSM_LIB_PATH") or os.path.join(_HERE, "libbayesssm_b200.so")   # override: A/B experiments only

# status codes / enums (include/bayesssm_b200.h)
OK, ERR_NEGATIVE_WEIGHT, ERR_ZERO_SUM, ERR_NAN_WEIGHT, ERR_BAD_ARG, ERR_PRIOR_INIT, ERR_CUDA, ERR_NVRTC, \
    ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CAPACITY, ERR_NCCL = range(12)
BPF, APF, RMPF = 0, 1, 2
SIS, SISR, SISAR = 0, 1, 2
STRATIFIED, SYSTEMATIC, MULTINOMIAL = 0, 1, 2
F32, F64 = 0, 1
ENGINE_AUTO, ENGINE_GENERAL, ENGINE_PERSISTENT, ENGINE_STREAM = 0, 1, 2, 3
MODEL_AR_SIN, MODEL_LG, MODEL_RW_DRIFT, MODEL_SIR_CB, MODEL_AR_COS, MODEL_RW2D, MODEL_SIR_GILLESPIE = range(7)
PRIOR_FLAT, PRIOR_NORMAL, PRIOR_EXP, PRIOR_UNIF, PRIOR_HALFNORMAL = range(5)
TR_IDENTITY, TR_LOG, TR_LOGIT = range(3)

ALGORITHMS = {"BPF": BPF, "APF": APF, "RMPF": RMPF}
RESAMPLE_ALGORITHMS = {"SIS": SIS, "SISR": SISR, "SISAR": SISAR}
RESAMPLE_FNS = {"stratified": STRATIFIED, "systematic": SYSTEMATIC, "multinomial": MULTINOMIAL}

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int32_p = C.POINTER(C.c_int32)


class NoiseBuffers(C.Structure):
    _fields_ = [(n, c_double_p) for n in (
        "z_init", "u_init", "z_trans", "u_trans", "z_trans2", "u_trans2",
        "u_resample", "u_resample_aux", "z_move", "u_move")]


class FilterConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("algorithm", C.c_int), ("resample_algorithm", C.c_int), ("resample_fn", C.c_int),
        ("threshold", C.c_double),
        ("num_particles", C.c_int), ("num_obs", C.c_int), ("dy", C.c_int),
        ("obs_times", c_int_p),
        ("num_filters", C.c_int), ("precision", C.c_int),
        ("seed", C.c_uint64), ("run_id", C.c_uint32), ("stream_base", C.c_uint32),
        ("noise", C.POINTER(NoiseBuffers)),
        ("return_particles", C.c_int), ("exact_resampling", C.c_int), ("engine", C.c_int), ("carry_weights", C.c_int),
    ]


class FilterResult(C.Structure):
    _fields_ = [
        ("loglike", c_double_p), ("loglike_history", c_double_p), ("ess", c_double_p), ("state_est", c_double_p),
        ("particles_history", c_double_p), ("weights_history", c_double_p),
        ("status", c_int32_p), ("early_exit", c_int32_p), ("n_resampled", c_int32_p),
        ("ancestors_history", c_int32_p), ("ancestors_aux_history", c_int32_p),
        ("kernel_ms", C.c_float),
    ]


class PmmhConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("algorithm", C.c_int), ("p", C.c_int),
        ("prior_kind", c_int_p), ("prior_a", c_double_p), ("prior_b", c_double_p),
        ("transform", c_int_p), ("pilot_proposal_sd", c_double_p),
        ("pilot_n", C.c_int), ("pilot_m", C.c_int), ("pilot_reps", C.c_int),
        ("pilot_resample_algorithm", C.c_int), ("pilot_resample_fn", C.c_int),
        ("m", C.c_int), ("num_chains", C.c_int), ("chain_id_base", C.c_uint32),
        ("fixed_num_particles", C.c_int),
        ("num_obs", C.c_int), ("dy", C.c_int), ("obs_times", c_int_p),
        ("consts", c_double_p), ("nconst", C.c_int),
        ("precision", C.c_int), ("seed", C.c_uint64),
        ("skip_pilot", C.c_int), ("proposal_chol_in", c_double_p),
        ("engine", C.c_int), ("return_latent_state_est", C.c_int),
    ]


class PmmhResult(C.Structure):
    _fields_ = [
        ("pilot_theta_chain", c_double_p), ("pilot_loglike_chain", c_double_p),
        ("pilot_theta_mean", c_double_p), ("pilot_theta_cov", c_double_p), ("pilot_loglikes", c_double_p),
        ("target_n", c_int32_p), ("proposal_chol", c_double_p),
        ("theta_chain", c_double_p), ("loglike_chain", c_double_p),
        ("n_accept", c_int32_p), ("status", c_int32_p),
        ("pilot_ms", C.c_float), ("main_ms", C.c_float),
        ("latent_state_chain", c_double_p),
        ("main_resampled_fraction", C.c_double),
    ]


# every symbol include/bayesssm_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "bssm_abi_version": (C.c_int, []),
    "bssm_last_error": (C.c_char_p, []),
    "bssm_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "bssm_destroy": (None, [_vp]),
    "bssm_device_info": (C.c_int, [_vp, C.c_char_p, c_int_p, c_int_p, c_int_p, C.POINTER(C.c_size_t)]),
    "bssm_launch_count": (C.c_int64, [_vp]),
    "bssm_synchronize": (C.c_int, [_vp]),
    "bssm_timer_start": (C.c_int, [_vp]),
    "bssm_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "bssm_resample_stratified": (C.c_int, [_vp, C.c_int, c_double_p, c_double_p, c_int32_p]),
    "bssm_resample_systematic": (C.c_int, [_vp, C.c_int, c_double_p, C.c_double, c_int32_p]),
    "bssm_resample_multinomial": (C.c_int, [_vp, C.c_int, c_double_p, c_double_p, c_int32_p]),
    "bssm_resample_cdf": (C.c_int, [_vp, C.c_int, c_double_p, c_double_p, c_double_p, C.POINTER(C.c_int64)]),
    "bssm_resample_device": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "bssm_dev_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "bssm_dev_free": (C.c_int, [_vp, _vp]),
    "bssm_dev_upload": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "bssm_dev_download": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "bssm_dev_fill_uniform": (C.c_int, [_vp, _vp, C.c_size_t, C.c_uint64]),
    "bssm_model_dims": (C.c_int, [_vp, C.c_int, c_int_p, c_int_p, c_int_p]),
    "bssm_model_noise_dims": (C.c_int, [_vp, C.c_int] + [c_int_p] * 6),
    "bssm_filter_run": (C.c_int, [_vp, C.POINTER(FilterConfig), c_double_p, c_double_p, C.POINTER(FilterResult)]),
    "bssm_filter_run_device": (C.c_int, [_vp, C.POINTER(FilterConfig), _vp, _vp, _vp, C.POINTER(C.c_float)]),
    "bssm_shard_unique_id": (C.c_int, [C.c_char_p, _vp]),
    "bssm_shard_init": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int, _vp]),
    "bssm_shard_finalize": (C.c_int, [_vp]),
    "bssm_shard_peer_export": (C.c_int, [_vp, _vp]),
    "bssm_shard_peer_attach": (C.c_int, [_vp, _vp]),
    "bssm_shard_peer_detach": (C.c_int, [_vp]),
    "bssm_shard_peer_active": (C.c_int, [_vp]),
    "bssm_shard_partition": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), c_int_p]),
    "bssm_filter_run_sharded": (C.c_int, [_vp, C.POINTER(FilterConfig), c_double_p, c_double_p, C.c_double,
                                          C.POINTER(FilterResult), c_int_p]),
    "bssm_model_compile": (C.c_int, [_vp, C.c_char_p, c_int_p]),
    "bssm_model_compile_log": (C.c_char_p, [_vp]),
    "bssm_pmmh_run": (C.c_int, [_vp, C.POINTER(PmmhConfig), c_double_p, c_double_p, C.POINTER(PmmhResult)]),
    "bssm_mcmc_diagnostics": (C.c_int, [_vp, c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p,
                                        c_int32_p, C.POINTER(C.c_float)]),
    "bssm_transform": (C.c_double, [C.c_double, C.c_int]),
    "bssm_back_transform": (C.c_double, [C.c_double, C.c_int]),
    "bssm_log_jacobian": (C.c_double, [c_double_p, c_int_p, C.c_int]),
    "bssm_log_prior": (C.c_double, [C.c_int, C.c_double, C.c_double, C.c_double]),
}


class EngineError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


_lib = None


def load_library():
    """Load the native engine; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA engine first "
            "(python -c 'import __graft_entry__ as g; g.build()' or python -m bayesssm_b200.build). "
            "bayesssm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the binding drift apart
        fn.restype = res
        fn.argtypes = args
    if lib.bssm_abi_version() != 1:
        raise ImportError("libbayesssm_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load_library().bssm_last_error().decode("utf-8", "replace")


def check(status: int):
    if status != OK:
        raise EngineError(status, last_error())


class Context:
    """Opaque engine context (one per process per GPU): stream, scratch HBM, compiled models."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.handle = _vp()
        check(self.lib.bssm_create(int(device), C.byref(self.handle)))
        self.device = device

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.bssm_destroy(self.handle)
            self.handle = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        name = C.create_string_buffer(256)
        sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
        mem = C.c_size_t()
        check(self.lib.bssm_device_info(self.handle, name, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem)))
        return {"name": name.value.decode(), "sm_count": sm.value, "cc": (maj.value, mnr.value), "global_mem": mem.value}

    def launch_count(self) -> int:
        return int(self.lib.bssm_launch_count(self.handle))

    def synchronize(self):
        check(self.lib.bssm_synchronize(self.handle))


_default_ctx = {}


def default_context(device: int | None = None) -> Context:
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("BSSM_USE_LOCAL_RANK") else 0
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = Context(device)
        _default_ctx[device] = ctx
    return ctx
