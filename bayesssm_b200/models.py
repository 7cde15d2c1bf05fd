"""Device models: what replaces the R closures init_fn / transition_fn / log_likelihood_fn /
aux_log_likelihood_fn / move_fn (R/particle_filter-doc.R:11-18).

A model is either built in (compiled into the engine) or a CUDA device-function snippet that the
engine compiles with NVRTC for sm_100a (`cuda_model`).  The filter front-ends take the model's slot
handles in the argument positions where the reference takes closures; plain Python callables are
rejected -- the engine runs no host closure per step and has no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field

from . import _native as nat


@dataclass(frozen=True)
class DeviceFn:
    """Handle of one operator slot of a device model (stands where the R closure stood)."""
    model: "DeviceModel"
    slot: str

    def __call__(self, *a, **k):
        raise TypeError(f"{self.slot} of model '{self.model.name}' is a device function; it runs inside the CUDA "
                        "engine and cannot be called from Python")


@dataclass(frozen=True)
class DeviceModel:
    name: str
    model_id: int
    param_names: tuple
    const_names: tuple = ()
    dim: int = 1
    has_aux: bool = True
    has_move: bool = True
    source: str | None = field(default=None, compare=False)

    @property
    def init_fn(self): return DeviceFn(self, "init_fn")
    @property
    def transition_fn(self): return DeviceFn(self, "transition_fn")
    @property
    def log_likelihood_fn(self): return DeviceFn(self, "log_likelihood_fn")
    @property
    def aux_log_likelihood_fn(self): return DeviceFn(self, "aux_log_likelihood_fn")
    @property
    def move_fn(self): return DeviceFn(self, "move_fn")


def nonlinear_ar():
    """README.md:137-146: x0~N(0,1); x_t = phi x + sin x + sigma_x v; y_t ~ N(x_t, sigma_y^2)."""
    return DeviceModel("nonlinear_ar", nat.MODEL_AR_SIN, ("phi", "sigma_x", "sigma_y"))


def nonlinear_ar_cos_obs():
    """R/pmmh.R:157-159: same dynamics, y_t ~ N(cos x_t, sigma_y^2)."""
    return DeviceModel("nonlinear_ar_cos_obs", nat.MODEL_AR_COS, ("phi", "sigma_x", "sigma_y"))


def linear_gaussian():
    """tests/testthat/test-pmmh_tuning.R:163-173: x_t = phi x + sigma_x v; y_t ~ N(x_t, sigma_y^2)."""
    return DeviceModel("linear_gaussian", nat.MODEL_LG, ("phi", "sigma_x", "sigma_y"))


def random_walk_drift():
    """tests/testthat/test-auxiliary_filter.R:17-27: x_t = x + N(mu, 1); y_t ~ N(x_t, sigma^2)."""
    return DeviceModel("random_walk_drift", nat.MODEL_RW_DRIFT, ("mu", "sigma"))


def sir_chain_binomial():
    """Chain-binomial stochastic SIR with Poisson observation of I."""
    return DeviceModel("sir_chain_binomial", nat.MODEL_SIR_CB, ("lambda", "gamma"), ("pop", "I0"), dim=2)


def sir_gillespie():
    """The same SIR model with the exact (Gillespie) daily step of the reference's vignette
    (vignettes/articles/stochastic-sir-model.Rmd:152-176, `epidemic_step`): a data-dependent number of uniforms per
    transition, drawn on demand from the particle's Philox stream."""
    return DeviceModel("sir_gillespie", nat.MODEL_SIR_GILLESPIE, ("lambda", "gamma"), ("pop", "I0"), dim=2)


def random_walk_2d():
    """tests/testthat/test-bootstrap_filter.R:211-217: 2-D random walk, flat likelihood."""
    return DeviceModel("random_walk_2d", nat.MODEL_RW2D, ("phi",), dim=2, has_aux=False, has_move=False)


BUILTIN = {m().name: m for m in (nonlinear_ar, nonlinear_ar_cos_obs, linear_gaussian, random_walk_drift,
                                 sir_chain_binomial, sir_gillespie, random_walk_2d)}


def resolve_model(*fns) -> DeviceModel:
    """All operator slots must be DeviceFn handles of ONE model (R closures cannot run on the GPU)."""
    model = None
    for fn in fns:
        if fn is None:
            continue
        if not isinstance(fn, DeviceFn):
            raise TypeError("init_fn / transition_fn / log_likelihood_fn must be device-model slots "
                            "(bayesssm_b200.models.*) or an NVRTC model; host closures are not supported "
                            "and there is no CPU fallback")
        if model is None:
            model = fn.model
        elif fn.model != model:
            raise ValueError("init_fn, transition_fn and log_likelihood_fn belong to different device models")
    if model is None:
        raise TypeError("no device model given")
    return model


def cuda_model(name, source, param_names, const_names=(), dim=1, has_aux=False, has_move=False, ctx=None):
    """A user model written as a CUDA device-function snippet (what the reference expresses as R closures,
    R/particle_filter-doc.R:11-18).  `source` defines `struct UserModel` following the contract at the top of
    bayesssm_b200/csrc/bssm_models.cuh; the engine compiles it with NVRTC for sm_100a together with its
    model-dependent kernels.  Needs a CUDA device (compilation loads the module)."""
    import ctypes as C
    ctx = ctx or nat.default_context()
    mid = C.c_int()
    st = ctx.lib.bssm_model_compile(ctx.handle, source.encode("utf-8"), C.byref(mid))
    if st != nat.OK:
        raise nat.EngineError(st, nat.last_error())
    d, nth, nc = C.c_int(), C.c_int(), C.c_int()
    nat.check(ctx.lib.bssm_model_dims(ctx.handle, mid.value, C.byref(d), C.byref(nth), C.byref(nc)))
    if (d.value, nth.value, nc.value) != (dim, len(param_names), len(const_names)):
        raise ValueError(f"UserModel declares D={d.value}, NTHETA={nth.value}, NCONST={nc.value}; "
                         f"the Python side says dim={dim}, {len(param_names)} parameters, {len(const_names)} constants")
    return DeviceModel(name, mid.value, tuple(param_names), tuple(const_names), dim=dim, has_aux=has_aux,
                       has_move=has_move, source=source)
