"""MCMC diagnostics called on every pmmh() return (R/ESS.R:30-104, R/rhat.R:27-67); host numpy
(post-hoc, not on the device hot path)."""
from __future__ import annotations

import warnings

import numpy as np


def _as_matrix(chains):
    if hasattr(chains, "to_numpy"):
        chains = chains.to_numpy()
    if not isinstance(chains, np.ndarray) or chains.ndim != 2:
        raise TypeError("Input 'chains' must be a matrix or a data frame.")
    return np.asarray(chains, dtype=np.float64)


def _acf(x):
    """stats::acf(x, lag.max = m - 1)$acf: biased autocovariance normalised by lag 0 (FFT)."""
    m = len(x)
    xc = x - x.mean()
    nfft = 1 << int(np.ceil(np.log2(2 * m)))
    f = np.fft.rfft(xc, nfft)
    ac = np.fft.irfft(f * np.conj(f), nfft)[:m] / m
    return ac / ac[0]


def ess(chains):
    """Multi-chain effective sample size, Geyer initial monotone sequence (R/ESS.R:30-104)."""
    mat = _as_matrix(chains)
    m, k = mat.shape
    if m < 2:
        raise ValueError("Number of iterations must be at least 2.")
    if k < 2:
        raise ValueError("Number of chains must be at least 2.")
    chain_means = mat.mean(axis=0)
    b = m / (k - 1) * np.sum((chain_means - chain_means.mean()) ** 2)
    chain_vars = mat.var(axis=0, ddof=1)
    if np.any(chain_vars == 0):
        warnings.warn("One or more chains have zero variance.")
        return float("nan")
    w = chain_vars.mean()
    var_hat = ((m - 1) / m) * w + b / m
    acf_matrix = np.stack([_acf(mat[:, i]) for i in range(k)], axis=1)  # [m][k]
    hat_rho = 1.0 - (w - (acf_matrix * chain_vars).sum(axis=1) / k) / var_hat
    max_pairs = (m - 1) // 2
    pairs = hat_rho[1:2 * max_pairs:2] + hat_rho[2:2 * max_pairs + 1:2]
    pairs = np.minimum.accumulate(pairs) if len(pairs) >= 2 else pairs
    neg = np.flatnonzero(pairs < 0)
    stop = neg[0] if len(neg) else len(pairs)
    tau = 1.0 + 2.0 * pairs[:stop].sum()
    return float(k * m / tau)


def rhat(chains):
    """Split-R-hat with the reference's [0.99, 1] -> 1 clamp (R/rhat.R:27-67)."""
    mat = _as_matrix(chains)
    m, k = mat.shape
    if m < 2:
        raise ValueError("Number of iterations must be at least 2.")
    if m % 2 == 1:
        mat = mat[:-1]
        m -= 1
    h = m // 2
    split = np.empty((h, 2 * k))
    split[:, 0::2] = mat[:h]
    split[:, 1::2] = mat[h:]
    chain_means = split.mean(axis=0)
    b = m / (2 * k - 1) * np.sum((chain_means - chain_means.mean()) ** 2)
    chain_vars = split.var(axis=0, ddof=1)
    if np.any(chain_vars == 0):
        warnings.warn("One or more chains have zero variance.")
        return float("nan")
    w = chain_vars.mean()
    var_hat = ((m - 1) / m) * w + b / m
    r = float(np.sqrt(var_hat / w))
    return 1.0 if 0.99 <= r <= 1.0 else r
