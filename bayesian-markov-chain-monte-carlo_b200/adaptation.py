"""Pooled adaptive-Metropolis proposal from sufficient statistics (extension).

The reference's own update (MCMC.py:162-204) is per chain, uses the last 10
samples and is either dead (list priors, q2) or dimensionally wrong (dict priors,
q3); it is reproduced on the device by ``RSFM_ADAPT_COMPAT``.  This module is the
corrected generalisation the north star asks for (Haario et al. 2001): the
covariance of ALL chains' draws, pooled over ranks by one all-reduce of
``[n, sum q (d), sum q q^T (lower triangle, row-major)]``, scaled by 2.38^2/d.
"""
import numpy as np

EPS_REL = 1e-10


def moments_from_suffstats(s, d):
    """(n, mean [d], cov [d, d]) with ddof = 1 from the packed sums."""
    s = np.asarray(s, dtype=np.float64)
    n = s[0]
    mean = s[1:1 + d] / n
    sec = np.zeros((d, d))
    sec[np.tril_indices(d)] = s[1 + d:1 + d + d * (d + 1) // 2]
    sec = sec + np.tril(sec, -1).T
    cov = (sec - n * np.outer(mean, mean)) / (n - 1.0)
    return n, mean, cov


def proposal_from_suffstats(s, d):
    """Packed proposal factor for ``rsfm_set_proposal_chol`` or None when not positive definite.

    d = 1: the proposal VARIANCE (the reference keeps a variance, MCMC.py:497);
    d = 3: row-major lower Cholesky factor [l00, l10, l11, l20, l21, l22].
    """
    n, mean, cov = moments_from_suffstats(s, d)
    v = (2.38 ** 2 / d) * cov
    if not np.all(np.isfinite(v)) or np.any(np.diag(v) <= 0.0):
        return None                      # degenerate moments: keep the current proposal (cf. quirk q4)
    v = v + np.diag(EPS_REL * np.diag(v))
    if d == 1:
        return np.array([v[0, 0]]) if v[0, 0] > 0 else None
    try:
        low = np.linalg.cholesky(v)
    except np.linalg.LinAlgError:
        return None
    return low[np.tril_indices(d)].copy()
