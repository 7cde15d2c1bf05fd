"""ORACLE -- test infrastructure only (never imported by the product path).

numpy restatement of the reference's MCMC diagnostics for one parameter given as an m x k matrix (iterations x
chains): ess() R/ESS.R:30-104 and rhat() R/rhat.R:27-67.  stats::acf is third-party to the reference (R itself,
unpinned): restated from its published definition -- x centred by its mean, acov[lag] = sum_i x[i] x[i+lag] / m,
acf = acov / acov[0] (src/library/stats/src/filter.c).  Pinned by the reference's own tests for these functions
(tests/testthat/test-ESS.R, test-rhat.R; statistical and error-string checks, see tests/test_diag_host.py); the
reference holds no known-answer vectors for them, so digit-level parity is unpinned.
"""
import numpy as np


def acf(x):
    """stats::acf(x, lag.max = m - 1, plot = FALSE)$acf: direct lagged sums."""
    m = len(x)
    xc = x - x.mean()
    ac = np.correlate(xc, xc, mode="full")[m - 1:] / m
    return ac / ac[0]


def ess_matrix(mat):
    """compute_ess_matrix (R/ESS.R:32-103).  Returns NaN where R returns NA (a chain with zero variance)."""
    mat = np.asarray(mat, dtype=np.float64)
    m, k = mat.shape
    if m < 2:
        raise ValueError("Number of iterations must be at least 2.")
    if k < 2:
        raise ValueError("Number of chains must be at least 2.")
    chain_means = mat.mean(axis=0)
    b = m / (k - 1) * np.sum((chain_means - chain_means.mean()) ** 2)
    chain_vars = mat.var(axis=0, ddof=1)
    if np.any(chain_vars == 0):
        return float("nan")
    w = chain_vars.mean()
    var_hat = ((m - 1) / m) * w + b / m
    acf_matrix = np.stack([acf(mat[:, i]) for i in range(k)], axis=1)
    hat_rho = np.array([1.0 - (w - (1.0 / k) * np.sum(chain_vars * acf_matrix[t])) / var_hat for t in range(m)])
    max_pairs = (m - 1) // 2
    pairs = [hat_rho[2 * t - 1] + hat_rho[2 * t] for t in range(1, max_pairs + 1)]
    for t in range(1, len(pairs)):
        if pairs[t] > pairs[t - 1]:
            pairs[t] = pairs[t - 1]
    sum_rho = 0.0
    for pr in pairs:
        if pr < 0:
            break
        sum_rho += pr
    return float(k * m / (1.0 + 2.0 * sum_rho))


def rhat_matrix(mat):
    """compute_rhat_matrix (R/rhat.R:28-67)."""
    mat = np.asarray(mat, dtype=np.float64)
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
        return float("nan")
    w = chain_vars.mean()
    var_hat = ((m - 1) / m) * w + b / m
    r = float(np.sqrt(var_hat / w))
    return 1.0 if 0.99 <= r <= 1.0 else r
