"""bootstrap_filter / auxiliary_filter / resample_move_filter with the reference's arguments and
return objects (R/bootstrap_filter.R:129-171, R/auxiliary_filter.R:163-216,
R/resample_move_filter.R:190-236, R/particle_filter_core.R:19-267), on the CUDA engine."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat
from .models import resolve_model


def _match_arg(value, choices, name):
    if isinstance(value, (tuple, list)):
        value = value[0]  # R's match.arg default: first choice
    if value not in choices:
        raise ValueError(f"'{name}' should be one of {', '.join(repr(c) for c in choices)}")
    return value


def _precision(precision):
    if precision in ("f64", "double", "fp64"):
        return nat.F64
    if precision in ("f32", "float", "fp32"):
        return nat.F32
    if precision in (nat.F32, nat.F64) and not isinstance(precision, str):
        return int(precision)
    raise ValueError("precision must be 'f64' or 'f32'")


def _theta_vector(model, params):
    names = model.param_names + model.const_names
    missing = [n for n in names if n not in params]
    if missing:
        raise ValueError(f"missing model parameter(s) {missing} for device model '{model.name}'")
    extra = [k for k in params if k not in names]
    if extra:
        raise ValueError(f"unused argument(s) {extra}: device model '{model.name}' takes {list(names)}")
    return np.array([float(params[n]) for n in names], dtype=np.float64)


def _particle_filter_core(y, num_particles, model, algorithm, resample_algorithm, resample_fn, threshold,
                          return_particles, obs_times, params, precision, seed, ctx, num_filters=1,
                          exact_resampling=-1, engine=nat.ENGINE_AUTO, run_id=0, stream_base=0, carry_weights=False):
    # validation as R/particle_filter_core.R:33-73
    if not (isinstance(num_particles, (int, np.integer)) and num_particles > 0):
        raise ValueError("Assertion on 'num_particles' failed: Must be a positive count")
    y = np.asarray(y, dtype=np.float64)
    if y.size == 0 or np.isnan(y).any():
        raise ValueError("Assertion on 'y' failed: Must be numeric without missing values")
    if y.ndim == 1:
        y = y[:, None]
    y = np.ascontiguousarray(y)
    T, dy = y.shape
    ot = None
    if obs_times is not None:
        ot_in = np.asarray(obs_times)
        if ot_in.shape != (T,):
            raise ValueError(f"Assertion on 'obs_times' failed: Must have length {T}")
        if not np.all(np.equal(np.mod(ot_in, 1), 0)):
            raise ValueError("Assertion on 'obs_times' failed: Must be of type 'integerish'")
        if np.any(ot_in < 1) or np.any(np.diff(ot_in) < 0):
            raise ValueError("Assertion on 'obs_times' failed: Must be sorted and >= 1")
        ot = np.ascontiguousarray(ot_in, dtype=np.int32)
    ctx = ctx or nat.default_context()
    theta = _theta_vector(model, params)
    theta = np.ascontiguousarray(np.broadcast_to(theta, (num_filters, len(theta))))
    d = model.dim
    cfg = nat.FilterConfig()
    cfg.model, cfg.algorithm = model.model_id, nat.ALGORITHMS[algorithm]
    cfg.resample_algorithm, cfg.resample_fn = nat.RESAMPLE_ALGORITHMS[resample_algorithm], nat.RESAMPLE_FNS[resample_fn]
    cfg.threshold = -1.0 if threshold is None else float(threshold)
    cfg.num_particles, cfg.num_obs, cfg.dy = int(num_particles), T, dy
    if ot is not None:
        cfg.obs_times = ot.ctypes.data_as(nat.c_int_p)
    cfg.num_filters = num_filters
    cfg.precision = _precision(precision)
    if seed is None:
        seed = int(np.random.default_rng().integers(0, 2**31 - 1))
    cfg.seed, cfg.run_id, cfg.stream_base = int(seed), run_id, stream_base
    cfg.return_particles = int(bool(return_particles))
    cfg.exact_resampling = exact_resampling
    cfg.engine = engine
    cfg.carry_weights = int(bool(carry_weights))
    N = int(num_particles)
    out = {"loglike": np.zeros(num_filters), "loglike_history": np.zeros((num_filters, T)),
           "ess": np.zeros((num_filters, T + 1)), "state_est": np.zeros((num_filters, T + 1, d)),
           "status": np.zeros(num_filters, dtype=np.int32), "early_exit": np.zeros(num_filters, dtype=np.int32),
           "n_resampled": np.zeros(num_filters, dtype=np.int32)}
    res = nat.FilterResult()
    for k in ("loglike", "loglike_history", "ess", "state_est"):
        setattr(res, k, out[k].ctypes.data_as(nat.c_double_p))
    for k in ("status", "early_exit", "n_resampled"):
        setattr(res, k, out[k].ctypes.data_as(nat.c_int32_p))
    if return_particles:
        out["particles_history"] = np.zeros((num_filters, T + 1, d, N))
        out["weights_history"] = np.zeros((num_filters, T + 1, N))
        res.particles_history = out["particles_history"].ctypes.data_as(nat.c_double_p)
        res.weights_history = out["weights_history"].ctypes.data_as(nat.c_double_p)
    nat.check(ctx.lib.bssm_filter_run(ctx.handle, C.byref(cfg), y.ctypes.data_as(nat.c_double_p),
                                      theta.ctypes.data_as(nat.c_double_p), C.byref(res)))
    if (out["status"] == nat.ERR_NAN_WEIGHT).any():
        raise ValueError("missing value where TRUE/FALSE needed")  # R/particle_filter_core.R:189 on NaN weights
    out["kernel_ms"] = res.kernel_ms
    return out


def _as_result(out, c, algorithm, resample_algorithm, return_particles, d):
    """One filter of the batch as the reference's result list (R/particle_filter_core.R:248-266)."""
    se = out["state_est"][c]
    r = {"state_est": se[:, 0].copy() if d == 1 else se.copy(),
         "ess": out["ess"][c].copy(), "loglike": float(out["loglike"][c]),
         "loglike_history": out["loglike_history"][c].copy(), "algorithm": algorithm}
    early = bool(out["early_exit"][c])
    if not early:
        r["resample_algorithm"] = resample_algorithm  # absent from the early-exit return (:189-202)
    if return_particles:
        ph = out["particles_history"][c]                     # [T+1][d][N]
        r["particles_history"] = ph.reshape(ph.shape[0], -1)  # as.numeric(matrix): column-major flatten
        r["weights_history"] = out["weights_history"][c].copy()
    r["n_resampled"] = int(out["n_resampled"][c])
    return r


def bootstrap_filter(y, num_particles, init_fn, transition_fn, log_likelihood_fn, obs_times=None,
                     resample_algorithm=("SISAR", "SISR", "SIS"), resample_fn=("stratified", "systematic", "multinomial"),
                     threshold=None, return_particles=True, *, precision="f64", seed=None, ctx=None,
                     engine=nat.ENGINE_AUTO, carry_weights=False, **params):
    """Bootstrap particle filter (R/bootstrap_filter.R:129-171).  Model parameters travel by name in **params.

    `carry_weights=True` (not in the reference, a stated deviation) carries the weights over steps that do not resample --
    standard SMC, where the reference's weights are the current likelihoods only; general kernels."""
    resample_algorithm = _match_arg(resample_algorithm, ("SISAR", "SISR", "SIS"), "resample_algorithm")
    resample_fn = _match_arg(resample_fn, ("stratified", "systematic", "multinomial"), "resample_fn")
    model = resolve_model(init_fn, transition_fn, log_likelihood_fn)
    out = _particle_filter_core(y, num_particles, model, "BPF", resample_algorithm, resample_fn, threshold,
                                return_particles, obs_times, params, precision, seed, ctx, engine=engine, carry_weights=carry_weights)
    return _as_result(out, 0, "BPF", resample_algorithm, return_particles, model.dim)


def auxiliary_filter(y, num_particles, init_fn, transition_fn, log_likelihood_fn, aux_log_likelihood_fn,
                     obs_times=None, resample_algorithm=("SISAR", "SISR", "SIS"),
                     resample_fn=("stratified", "systematic", "multinomial"), threshold=None, return_particles=True,
                     *, precision="f64", seed=None, ctx=None, **params):
    """Auxiliary particle filter (R/auxiliary_filter.R:163-216)."""
    resample_algorithm = _match_arg(resample_algorithm, ("SISAR", "SISR", "SIS"), "resample_algorithm")
    resample_fn = _match_arg(resample_fn, ("stratified", "systematic", "multinomial"), "resample_fn")
    model = resolve_model(init_fn, transition_fn, log_likelihood_fn, aux_log_likelihood_fn)
    if not model.has_aux:
        raise ValueError(f"device model '{model.name}' has no aux_log_likelihood_fn")
    out = _particle_filter_core(y, num_particles, model, "APF", resample_algorithm, resample_fn, threshold,
                                return_particles, obs_times, params, precision, seed, ctx)
    return _as_result(out, 0, "APF", resample_algorithm, return_particles, model.dim)


def resample_move_filter(y, num_particles, init_fn, transition_fn, log_likelihood_fn, move_fn, obs_times=None,
                         resample_fn=("stratified", "systematic", "multinomial"), threshold=None,
                         return_particles=True, *, precision="f64", seed=None, ctx=None, **params):
    """Resample-move particle filter (R/resample_move_filter.R:190-236): always SISR; a user
    resample_algorithm is dropped as in the reference (:213-216)."""
    params.pop("resample_algorithm", None)
    resample_fn = _match_arg(resample_fn, ("stratified", "systematic", "multinomial"), "resample_fn")
    model = resolve_model(init_fn, transition_fn, log_likelihood_fn, move_fn)
    if not model.has_move:
        raise ValueError(f"device model '{model.name}' has no move_fn")
    out = _particle_filter_core(y, num_particles, model, "RMPF", "SISR", resample_fn, threshold,
                                return_particles, obs_times, params, precision, seed, ctx)
    return _as_result(out, 0, "RMPF", "SISR", return_particles, model.dim)


def particle_filter(*args, **kwargs):
    """The generic name the north star lists; the reference's documentation stub (R/particle_filter.R)
    points at bootstrap_filter."""
    return bootstrap_filter(*args, **kwargs)


def batched_bootstrap_filter(y, num_particles, model, num_filters, resample_algorithm="SISAR",
                             resample_fn="stratified", threshold=None, *, precision="f32", seed=0, ctx=None,
                             engine=nat.ENGINE_AUTO, **params):
    """`num_filters` replicate bootstrap filters in one launch (config C3); returns the raw batch arrays."""
    return _particle_filter_core(y, num_particles, model, "BPF", resample_algorithm, resample_fn, threshold, False,
                                 None, params, precision, seed, ctx, num_filters=num_filters, engine=engine)
