"""MCMC diagnostics called on every pmmh() return: ess() (R/ESS.R:30-145) and rhat() (R/rhat.R:27-108), computed on
the device for all parameters at once through bssm_mcmc_diagnostics (csrc/bssm_diag.cuh).  This module only
checks the input the way the reference does and maps NA / warnings; nothing is computed on the CPU."""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _native as nat

_ZERO_VAR = "One or more chains have zero variance."


def device_diagnostics(draws, burn_in=0, want_ess=True, want_rhat=True, ctx=None):
    """draws [k][m][p] (chains, iterations, parameters) -> dict(ess [p], rhat [p], flags [p], device_ms)."""
    ctx = ctx or nat.default_context()
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    k, m, p = draws.shape
    out_ess, out_rhat = np.full(p, np.nan), np.full(p, np.nan)
    flags = np.zeros(p, dtype=np.int32)
    ms = C.c_float(0.0)
    st = ctx.lib.bssm_mcmc_diagnostics(ctx.handle, draws.ctypes.data_as(nat.c_double_p), k, m, p, int(burn_in),
                                       out_ess.ctypes.data_as(nat.c_double_p) if want_ess else None,
                                       out_rhat.ctypes.data_as(nat.c_double_p) if want_rhat else None,
                                       flags.ctypes.data_as(nat.c_int32_p), C.byref(ms))
    if st == nat.ERR_BAD_ARG:
        raise ValueError(nat.last_error())   # "Number of iterations / chains must be at least 2."
    nat.check(st)
    return {"ess": out_ess, "rhat": out_rhat, "flags": flags, "device_ms": ms.value}


def _frame_to_draws(df):
    """Data frame with a 'chain' column -> (parameter names, draws [k][m][p]) (R/ESS.R:114-141, R/rhat.R:77-104)."""
    if "chain" not in df.columns:
        raise ValueError("Data frame must contain a 'chain' column.")
    params = [c for c in df.columns if c != "chain"]
    ids = list(dict.fromkeys(df["chain"].tolist()))
    per_chain = [df.loc[df["chain"] == i, params].to_numpy(dtype=np.float64) for i in ids]
    if len({a.shape[0] for a in per_chain}) != 1:
        raise ValueError("Not all chains have the same number of iterations.")
    return params, np.stack(per_chain, axis=0)


def _run(chains, which, ctx):
    if hasattr(chains, "columns") and hasattr(chains, "loc"):
        params, draws = _frame_to_draws(chains)
    elif isinstance(chains, np.ndarray) and chains.ndim == 2:
        params, draws = None, np.asarray(chains, dtype=np.float64).T[:, :, None]   # m x k matrix -> [k][m][1]
    else:
        raise TypeError("Input must be a matrix or a data frame with a 'chain' column.")
    k, m, _ = draws.shape
    if m < 2:
        raise ValueError("Number of iterations must be at least 2.")
    if which == "ess" and k < 2:
        raise ValueError("Number of chains must be at least 2.")
    r = device_diagnostics(draws, 0, want_ess=which == "ess", want_rhat=which == "rhat", ctx=ctx)
    if np.any(r["flags"] & (1 if which == "ess" else 2)):
        warnings.warn(_ZERO_VAR)
    vals = r[which]
    return float(vals[0]) if params is None else {name: float(v) for name, v in zip(params, vals)}


def ess(chains, ctx=None):
    """Multi-chain effective sample size, Geyer's initial monotone sequence (R/ESS.R:30-145).  `chains`: an m x k
    matrix (iterations x chains) -> float, or a data frame with a 'chain' column -> dict per parameter."""
    return _run(chains, "ess", ctx)


def rhat(chains, ctx=None):
    """Split R-hat; values in [0.99, 1] are reported as 1 (R/rhat.R:27-108).  Input as for ess()."""
    return _run(chains, "rhat", ctx)
