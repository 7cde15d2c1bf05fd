"""The three registered .Call routines of the reference (R/RcppExports.R:4-14, src/resampling.cpp:5-66)
and their R wrappers (R/resampling.R:13-69), on the CUDA engine.  Uniforms are drawn on the host with the
caller's numpy Generator (the R shim uses unif_rand() so set.seed() keeps its meaning)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def _rng(rng):
    return rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)


def _open_unit(rng, n):
    u = rng.random(n)
    return np.where(u <= 0.0, np.finfo(float).tiny, u)  # R's unif_rand() is open on (0, 1)


def _call(kind, n, weights, u, ctx=None):
    ctx = ctx or nat.default_context()
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if w.ndim != 1 or len(w) != int(n):
        raise ValueError("Length of weights must match n")
    out = np.zeros(int(n), dtype=np.int32)
    wp, op = w.ctypes.data_as(nat.c_double_p), out.ctypes.data_as(nat.c_int32_p)
    if kind == "systematic":
        st = ctx.lib.bssm_resample_systematic(ctx.handle, int(n), wp, float(u), op)
    else:
        u = np.ascontiguousarray(u, dtype=np.float64)
        fn = ctx.lib.bssm_resample_stratified if kind == "stratified" else ctx.lib.bssm_resample_multinomial
        st = fn(ctx.handle, int(n), wp, u.ctypes.data_as(nat.c_double_p), op)
    if st:
        raise ValueError(nat.last_error())  # "Weights must be non-negative" / "Sum of weights must be greater than 0"
    return out


def resample_multinomial_cpp(n, weights, rng=None, u=None):
    """1-based ancestor indices (src/resampling.cpp:5-13).  Natural-order inverse-cdf draws: same law as
    Rcpp::sample, not the same index stream."""
    return _call("multinomial", n, weights, _open_unit(_rng(rng), int(n)) if u is None else u)


def resample_stratified_cpp(n, weights, rng=None, u=None):
    """1-based ancestor indices (src/resampling.cpp:16-40)."""
    return _call("stratified", n, weights, _open_unit(_rng(rng), int(n)) if u is None else u)


def resample_systematic_cpp(n, weights, rng=None, u=None):
    """1-based ancestor indices (src/resampling.cpp:43-66)."""
    return _call("systematic", n, weights, _open_unit(_rng(rng), 1)[0] if u is None else u)


def _resample(cpp, particles, weights, rng):
    particles = np.asarray(particles)
    n = particles.shape[0]
    if len(weights) != n:  # R/resampling.R:15-17
        raise ValueError("Number of particles must match the length of weights")
    idx = cpp(n, weights, rng=rng) - 1
    return particles[idx]  # vector or row-gathered matrix (R/resampling.R:20,40,60)


def resample_multinomial(particles, weights, rng=None): return _resample(resample_multinomial_cpp, particles, weights, rng)
def resample_stratified(particles, weights, rng=None): return _resample(resample_stratified_cpp, particles, weights, rng)
def resample_systematic(particles, weights, rng=None): return _resample(resample_systematic_cpp, particles, weights, rng)
