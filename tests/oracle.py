"""ctypes binding of the CPU oracle (oracle/pf_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "libpf_oracle.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
i32p = C.POINTER(C.c_int32)


def build_oracle():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("pf_oracle.c", "pf_oracle.h", "Makefile")]
    if os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB


class NoiseBuffers(C.Structure):
    _fields_ = [(n, dp) for n in ("z_init", "u_init", "z_trans", "u_trans", "z_trans2", "u_trans2",
                                  "u_resample", "u_resample_aux", "z_move", "u_move")]


class FilterConfig(C.Structure):
    _fields_ = [("model", C.c_int), ("algorithm", C.c_int), ("resample_algorithm", C.c_int), ("resample_fn", C.c_int),
                ("threshold", C.c_double), ("num_particles", C.c_int), ("num_obs", C.c_int), ("dy", C.c_int),
                ("obs_times", ip), ("return_particles", C.c_int), ("noise", C.POINTER(NoiseBuffers)),
                ("seed", C.c_uint64), ("run_id", C.c_uint32), ("stream", C.c_uint32), ("carry_weights", C.c_int)]


class FilterResult(C.Structure):
    _fields_ = [("state_est", dp), ("ess", dp), ("loglike", C.c_double), ("loglike_history", dp),
                ("particles_history", dp), ("weights_history", dp), ("ancestors_history", i32p),
                ("ancestors_aux_history", i32p), ("early_exit", C.c_int), ("n_resampled", C.c_int)]


class PmmhConfig(C.Structure):
    _fields_ = [("model", C.c_int), ("algorithm", C.c_int), ("p", C.c_int),
                ("prior_kind", ip), ("prior_a", dp), ("prior_b", dp), ("transform", ip),
                ("pilot_proposal_sd", dp), ("pilot_n", C.c_int), ("pilot_m", C.c_int), ("pilot_reps", C.c_int),
                ("pilot_resample_algorithm", C.c_int), ("pilot_resample_fn", C.c_int),
                ("m", C.c_int), ("burn_in", C.c_int), ("fixed_num_particles", C.c_int),
                ("num_obs", C.c_int), ("dy", C.c_int), ("obs_times", ip), ("seed", C.c_uint64),
                ("consts", dp), ("nconst", C.c_int)]


class PmmhChainResult(C.Structure):
    _fields_ = [("pilot_theta_chain", dp), ("pilot_loglike_chain", dp), ("pilot_theta_mean", dp),
                ("pilot_theta_cov", dp), ("pilot_loglikes", dp), ("target_n", C.c_int), ("proposal_chol", dp),
                ("theta_chain", dp), ("loglike_chain", dp), ("n_accept", C.c_int), ("latent_state_chain", dp)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(LIB)
        L.orc_resample_stratified.argtypes = [C.c_int, dp, dp, i32p]
        L.orc_resample_systematic.argtypes = [C.c_int, dp, C.c_double, i32p]
        L.orc_resample_multinomial_invcdf.argtypes = [C.c_int, dp, dp, i32p]
        L.orc_resample_multinomial_rcpp.argtypes = [C.c_int, dp, dp, i32p]
        L.orc_resample_cdf.argtypes = [C.c_int, dp, dp, dp]
        L.orc_noise_uniform.restype = C.c_double
        L.orc_noise_uniform.argtypes = [C.c_uint64] + [C.c_uint32] * 6
        L.orc_noise_normal.restype = C.c_double
        L.orc_noise_normal.argtypes = [C.c_uint64] + [C.c_uint32] * 6
        L.orc_model_dims.argtypes = [C.c_int] + [ip] * 9
        L.orc_particle_filter.argtypes = [C.POINTER(FilterConfig), dp, dp, C.POINTER(FilterResult)]
        L.orc_kalman_loglik.restype = C.c_double
        L.orc_kalman_loglik.argtypes = [C.c_int, dp, C.c_double, C.c_double, C.c_double]
        for f in ("orc_transform", "orc_back_transform"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double, C.c_int]
        L.orc_log_jacobian.restype = C.c_double
        L.orc_log_jacobian.argtypes = [dp, ip, C.c_int]
        L.orc_log_prior.restype = C.c_double
        L.orc_log_prior.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_pmmh_chain.argtypes = [C.POINTER(PmmhConfig), dp, dp, C.c_uint32, C.POINTER(PmmhChainResult)]
        L.orc_bench_bootstrap_filter.restype = C.c_double
        L.orc_bench_bootstrap_filter.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_int, C.c_int, C.c_double,
                                                 C.c_uint32, dp, ip]
        _lib = L
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(dp) if a is not None else None


ERRORS = {1: "Weights must be non-negative", 2: "Sum of weights must be greater than 0", 3: "NaN weight",
          4: "bad argument", 5: "Initial parameter values are invalid"}


class OracleError(ValueError):
    def __init__(self, status):
        super().__init__(ERRORS.get(status, f"status {status}"))
        self.status = status


def resample(kind: str, w, u):
    """kind in stratified/systematic/multinomial (inverse-cdf)/multinomial_rcpp -> 1-based int32 ancestors"""
    w = _d(w)
    n = len(w)
    out = np.zeros(n, dtype=np.int32)
    L = lib()
    if kind == "systematic":
        st = L.orc_resample_systematic(n, _p(w), float(np.ravel(u)[0]), out.ctypes.data_as(i32p))
    else:
        u = _d(u)
        fn = {"stratified": L.orc_resample_stratified, "multinomial": L.orc_resample_multinomial_invcdf,
              "multinomial_rcpp": L.orc_resample_multinomial_rcpp}[kind]
        st = fn(n, _p(w), _p(u), out.ctypes.data_as(i32p))
    if st:
        raise OracleError(st)
    return out


def cdf(w):
    w = _d(w)
    out = np.zeros(len(w))
    tot = C.c_double()
    st = lib().orc_resample_cdf(len(w), _p(w), _p(out), C.cast(C.byref(tot), dp))
    if st:
        raise OracleError(st)
    return out, tot.value


def model_dims(model: int):
    v = [C.c_int() for _ in range(9)]
    st = lib().orc_model_dims(model, *[C.byref(x) for x in v])
    if st:
        raise OracleError(st)
    keys = ("d", "ntheta", "nconst", "nz_init", "nu_init", "nz_trans", "nu_trans", "nz_move", "nu_move")
    return dict(zip(keys, [x.value for x in v]))


def make_noise(model: int, N: int, T: int, n_time: int, rng: np.random.Generator):
    """Random injected-noise buffers in the layout of orc_noise_buffers (dict of float64 arrays)."""
    md = model_dims(model)
    nb = {
        "z_init": rng.standard_normal((max(md["nz_init"], 1), N)),
        "u_init": rng.random((max(md["nu_init"], 1), N)),
        "z_trans": rng.standard_normal((n_time, max(md["nz_trans"], 1), N)),
        "u_trans": rng.random((n_time, max(md["nu_trans"], 1), N)),
        "z_trans2": rng.standard_normal((T, max(md["nz_trans"], 1), N)),
        "u_trans2": rng.random((T, max(md["nu_trans"], 1), N)),
        "u_resample": rng.random((T, N)),
        "u_resample_aux": rng.random((T, N)),
        "z_move": rng.standard_normal((T, max(md["nz_move"], 1), N)),
        "u_move": rng.random((T, max(md["nu_move"], 1), N)),
    }
    return {k: np.ascontiguousarray(v) for k, v in nb.items()}


def particle_filter(model, algorithm, resample_algorithm, resample_fn, N, y, theta, threshold=-1.0, obs_times=None,
                    noise=None, seed=0, run_id=0, stream=0, return_particles=False, want_ancestors=False, carry_weights=False):
    y = _d(y)
    if y.ndim == 1:
        y = y[:, None]
    T, dy = y.shape
    md = model_dims(model)
    d = md["d"]
    theta = _d(theta)
    cfg = FilterConfig()
    cfg.model, cfg.algorithm, cfg.resample_algorithm, cfg.resample_fn = model, algorithm, resample_algorithm, resample_fn
    cfg.threshold = threshold
    cfg.num_particles, cfg.num_obs, cfg.dy = N, T, dy
    ot = None
    if obs_times is not None:
        ot = np.ascontiguousarray(obs_times, dtype=np.int32)
        cfg.obs_times = ot.ctypes.data_as(ip)
    cfg.return_particles = int(return_particles)
    nbs = None
    if noise is not None:
        nbs = NoiseBuffers()
        for k, v in noise.items():
            setattr(nbs, k, _p(v))
        cfg.noise = C.pointer(nbs)
    cfg.seed, cfg.run_id, cfg.stream = seed, run_id, stream
    cfg.carry_weights = int(carry_weights)
    res = FilterResult()
    out = {"state_est": np.zeros((T + 1, d)), "ess": np.zeros(T + 1), "loglike_history": np.zeros(T)}
    res.state_est, res.ess, res.loglike_history = _p(out["state_est"]), _p(out["ess"]), _p(out["loglike_history"])
    if return_particles:
        out["particles_history"] = np.zeros((T + 1, d, N))
        out["weights_history"] = np.zeros((T + 1, N))
        res.particles_history, res.weights_history = _p(out["particles_history"]), _p(out["weights_history"])
    if want_ancestors:
        out["ancestors_history"] = np.zeros((T, N), dtype=np.int32)
        out["ancestors_aux_history"] = np.zeros((T, N), dtype=np.int32)
        res.ancestors_history = out["ancestors_history"].ctypes.data_as(i32p)
        res.ancestors_aux_history = out["ancestors_aux_history"].ctypes.data_as(i32p)
    st = lib().orc_particle_filter(C.byref(cfg), _p(y), _p(theta), C.byref(res))
    out["status"] = st
    out["loglike"] = res.loglike
    out["early_exit"] = res.early_exit
    out["n_resampled"] = res.n_resampled
    return out


def kalman_loglik(y, phi, sigma_x, sigma_y):
    y = _d(y).ravel()
    return lib().orc_kalman_loglik(len(y), _p(y), phi, sigma_x, sigma_y)


def pmmh_chain(model, algorithm, y, init_theta, prior_kind, prior_a, prior_b, transform, pilot_proposal_sd,
               pilot_n, pilot_m, pilot_reps, m, chain_id, seed, pilot_resample_algorithm=2, pilot_resample_fn=0,
               fixed_num_particles=0, consts=None, obs_times=None, return_latent_state_est=False):
    y = _d(y)
    if y.ndim == 1:
        y = y[:, None]
    T, dy = y.shape
    p = len(init_theta)
    cfg = PmmhConfig()
    pk = np.ascontiguousarray(prior_kind, dtype=np.int32)
    pa, pb = _d(prior_a), _d(prior_b)
    tr = np.ascontiguousarray(transform, dtype=np.int32)
    sd = _d(pilot_proposal_sd)
    cfg.model, cfg.algorithm, cfg.p = model, algorithm, p
    cfg.prior_kind, cfg.prior_a, cfg.prior_b = pk.ctypes.data_as(ip), _p(pa), _p(pb)
    cfg.transform, cfg.pilot_proposal_sd = tr.ctypes.data_as(ip), _p(sd)
    cfg.pilot_n, cfg.pilot_m, cfg.pilot_reps = pilot_n, pilot_m, pilot_reps
    cfg.pilot_resample_algorithm, cfg.pilot_resample_fn = pilot_resample_algorithm, pilot_resample_fn
    cfg.m, cfg.burn_in, cfg.fixed_num_particles = m, 0, fixed_num_particles
    cfg.num_obs, cfg.dy = T, dy
    ot = None
    if obs_times is not None:
        ot = np.ascontiguousarray(obs_times, dtype=np.int32)
        cfg.obs_times = ot.ctypes.data_as(ip)
    cfg.seed = seed
    cs = _d(consts) if consts is not None else np.zeros(1)
    cfg.consts, cfg.nconst = _p(cs), (len(consts) if consts is not None else 0)
    out = {"pilot_theta_chain": np.zeros((pilot_m, p)), "pilot_loglike_chain": np.zeros(pilot_m),
           "pilot_theta_mean": np.zeros(p), "pilot_theta_cov": np.zeros((p, p)), "pilot_loglikes": np.zeros(pilot_reps),
           "proposal_chol": np.zeros((p, p)), "theta_chain": np.zeros((m, p)), "loglike_chain": np.zeros(m)}
    res = PmmhChainResult()
    for k, v in out.items():
        setattr(res, k, _p(v))
    if return_latent_state_est:
        out["latent_state_chain"] = np.zeros((m, T + 1, model_dims(model)["d"]))
        res.latent_state_chain = _p(out["latent_state_chain"])
    it = _d(init_theta)
    st = lib().orc_pmmh_chain(C.byref(cfg), _p(y), _p(it), chain_id, C.byref(res))
    out["status"] = st
    out["target_n"] = res.target_n
    out["n_accept"] = res.n_accept
    return out


def transform(th, tr):
    return lib().orc_transform(float(th), int(tr))


def back_transform(z, tr):
    return lib().orc_back_transform(float(z), int(tr))


def log_jacobian(theta, tr):
    th = _d(theta)
    t = np.ascontiguousarray(tr, dtype=np.int32)
    return lib().orc_log_jacobian(_p(th), t.ctypes.data_as(ip), len(th))


def log_prior(kind, a, b, x):
    return lib().orc_log_prior(int(kind), float(a), float(b), float(x))


def bench_bootstrap_filter(model, N, y, theta, resample_algorithm=2, resample_fn=0, threshold=-1.0, seed=1405):
    y = _d(y).ravel()
    th = _d(theta)
    ll = C.c_double()
    nres = C.c_int()
    secs = lib().orc_bench_bootstrap_filter(model, N, len(y), _p(y), _p(th), resample_algorithm, resample_fn,
                                            threshold, seed, C.cast(C.byref(ll), dp), C.byref(nres))
    return secs, ll.value, nres.value
