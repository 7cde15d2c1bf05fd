"""Thin numpy helpers over the C ABI used by the GPU parity tests (they call through the C ABI,
exactly as the R .Call shim would)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from bayesssm_b200 import _native as nat

dp = nat.c_double_p
i32p = nat.c_int32_p


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(dp) if a is not None else None


def resample(ctx, kind, w, u):
    w = _d(w)
    n = len(w)
    out = np.zeros(n, dtype=np.int32)
    if kind == "systematic":
        st = ctx.lib.bssm_resample_systematic(ctx.handle, n, _p(w), float(np.ravel(u)[0]), out.ctypes.data_as(i32p))
    else:
        u = _d(u)
        fn = ctx.lib.bssm_resample_stratified if kind == "stratified" else ctx.lib.bssm_resample_multinomial
        st = fn(ctx.handle, n, _p(w), _p(u), out.ctypes.data_as(i32p))
    nat.check(st)
    return out


def cdf(ctx, w):
    w = _d(w)
    n = len(w)
    out = np.zeros(n)
    tot = C.c_double()
    ser = C.c_int64()
    nat.check(ctx.lib.bssm_resample_cdf(ctx.handle, n, _p(w), _p(out), C.cast(C.byref(tot), dp), C.byref(ser)))
    return out, tot.value, ser.value


def filter_run(ctx, model, algorithm, resample_algorithm, resample_fn, N, y, theta, threshold=-1.0, obs_times=None,
               noise=None, seed=0, run_id=0, stream_base=0, precision=nat.F64, return_particles=False,
               want_ancestors=False, exact=-1, engine=nat.ENGINE_AUTO, num_filters=None):
    y = _d(y)
    if y.ndim == 1:
        y = y[:, None]
    T, dy = y.shape
    theta = _d(theta)
    if theta.ndim == 1:
        theta = theta[None, :]
    Cn = num_filters or theta.shape[0]
    if theta.shape[0] != Cn:
        theta = np.ascontiguousarray(np.broadcast_to(theta, (Cn, theta.shape[1])))
    d, nth, nc = C.c_int(), C.c_int(), C.c_int()
    nat.check(ctx.lib.bssm_model_dims(ctx.handle, model, C.byref(d), C.byref(nth), C.byref(nc)))
    d = d.value
    cfg = nat.FilterConfig()
    cfg.model, cfg.algorithm, cfg.resample_algorithm, cfg.resample_fn = model, algorithm, resample_algorithm, resample_fn
    cfg.threshold = threshold
    cfg.num_particles, cfg.num_obs, cfg.dy = N, T, dy
    ot = None
    if obs_times is not None:
        ot = np.ascontiguousarray(obs_times, dtype=np.int32)
        cfg.obs_times = ot.ctypes.data_as(nat.c_int_p)
    cfg.num_filters, cfg.precision = Cn, precision
    cfg.seed, cfg.run_id, cfg.stream_base = seed, run_id, stream_base
    nbs = None
    if noise is not None:
        nbs = nat.NoiseBuffers()
        for k, v in noise.items():
            setattr(nbs, k, _p(v))
        cfg.noise = C.pointer(nbs)
    cfg.return_particles = int(return_particles)
    cfg.exact_resampling = exact
    cfg.engine = engine
    out = {"loglike": np.zeros(Cn), "loglike_history": np.zeros((Cn, T)), "ess": np.zeros((Cn, T + 1)),
           "state_est": np.zeros((Cn, T + 1, d)), "status": np.zeros(Cn, dtype=np.int32),
           "early_exit": np.zeros(Cn, dtype=np.int32), "n_resampled": np.zeros(Cn, dtype=np.int32)}
    res = nat.FilterResult()
    for k in ("loglike", "loglike_history", "ess", "state_est"):
        setattr(res, k, _p(out[k]))
    for k in ("status", "early_exit", "n_resampled"):
        setattr(res, k, out[k].ctypes.data_as(i32p))
    if return_particles:
        out["particles_history"] = np.zeros((Cn, T + 1, d, N))
        out["weights_history"] = np.zeros((Cn, T + 1, N))
        res.particles_history, res.weights_history = _p(out["particles_history"]), _p(out["weights_history"])
    if want_ancestors:
        out["ancestors_history"] = np.zeros((Cn, T, N), dtype=np.int32)
        out["ancestors_aux_history"] = np.zeros((Cn, T, N), dtype=np.int32)
        res.ancestors_history = out["ancestors_history"].ctypes.data_as(i32p)
        res.ancestors_aux_history = out["ancestors_aux_history"].ctypes.data_as(i32p)
    nat.check(ctx.lib.bssm_filter_run(ctx.handle, C.byref(cfg), _p(y), _p(theta), C.byref(res)))
    out["kernel_ms"] = res.kernel_ms
    return out
