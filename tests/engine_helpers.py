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
               want_ancestors=False, exact=-1, engine=nat.ENGINE_AUTO, num_filters=None, carry_weights=False):
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
    cfg.carry_weights = int(carry_weights)
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


def pmmh_run(ctx, model, algorithm, y, init_theta, prior_kind, prior_a, prior_b, transform, pilot_proposal_sd,
             pilot_n, pilot_m, pilot_reps, m, seed, chain_id_base=0, pilot_resample_algorithm=2, pilot_resample_fn=0,
             fixed_num_particles=0, consts=None, obs_times=None, precision=nat.F64, skip_pilot=False,
             proposal_chol=None, engine=nat.ENGINE_AUTO, return_latent_state_est=False, state_dim=1):
    y = _d(y)
    if y.ndim == 1:
        y = y[:, None]
    T, dy = y.shape
    init_theta = _d(init_theta)
    if init_theta.ndim == 1:
        init_theta = init_theta[None, :]
    Cn, p = init_theta.shape
    cfg = nat.PmmhConfig()
    pk = np.ascontiguousarray(prior_kind, dtype=np.int32)
    pa, pb = _d(prior_a), _d(prior_b)
    tr = np.ascontiguousarray(transform, dtype=np.int32)
    sd = _d(pilot_proposal_sd)
    cfg.model, cfg.algorithm, cfg.p = model, algorithm, p
    cfg.prior_kind, cfg.prior_a, cfg.prior_b = pk.ctypes.data_as(nat.c_int_p), _p(pa), _p(pb)
    cfg.transform, cfg.pilot_proposal_sd = tr.ctypes.data_as(nat.c_int_p), _p(sd)
    cfg.pilot_n, cfg.pilot_m, cfg.pilot_reps = pilot_n, pilot_m, pilot_reps
    cfg.pilot_resample_algorithm, cfg.pilot_resample_fn = pilot_resample_algorithm, pilot_resample_fn
    cfg.m, cfg.num_chains, cfg.chain_id_base = m, Cn, chain_id_base
    cfg.fixed_num_particles = fixed_num_particles
    cfg.num_obs, cfg.dy = T, dy
    ot = None
    if obs_times is not None:
        ot = np.ascontiguousarray(obs_times, dtype=np.int32)
        cfg.obs_times = ot.ctypes.data_as(nat.c_int_p)
    cs = _d(consts) if consts is not None else np.zeros(1)
    cfg.consts, cfg.nconst = _p(cs), (len(consts) if consts is not None else 0)
    cfg.precision, cfg.seed = precision, seed
    cfg.skip_pilot = int(skip_pilot)
    chol_in = None
    if proposal_chol is not None:
        chol_in = _d(proposal_chol)
        cfg.proposal_chol_in = _p(chol_in)
    cfg.engine = engine
    pm = 1 if skip_pilot else pilot_m
    reps = 1 if skip_pilot else pilot_reps
    out = {"pilot_theta_chain": np.zeros((Cn, pm, p)), "pilot_loglike_chain": np.zeros((Cn, pm)),
           "pilot_theta_mean": np.zeros((Cn, p)), "pilot_theta_cov": np.zeros((Cn, p, p)),
           "pilot_loglikes": np.zeros((Cn, reps)), "proposal_chol": np.zeros((Cn, p, p)),
           "theta_chain": np.zeros((Cn, m, p)), "loglike_chain": np.zeros((Cn, m))}
    res = nat.PmmhResult()
    for k, v in out.items():
        setattr(res, k, _p(v))
    for k in ("target_n", "n_accept", "status"):
        out[k] = np.zeros(Cn, dtype=np.int32)
        setattr(res, k, out[k].ctypes.data_as(i32p))
    if return_latent_state_est:
        cfg.return_latent_state_est = 1
        out["latent_state_chain"] = np.zeros((Cn, m, T + 1, state_dim))
        res.latent_state_chain = _p(out["latent_state_chain"])
    nat.check(ctx.lib.bssm_pmmh_run(ctx.handle, C.byref(cfg), _p(y), _p(init_theta), C.byref(res)))
    out["pilot_ms"], out["main_ms"] = res.pilot_ms, res.main_ms
    return out
