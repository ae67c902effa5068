"""Convergence diagnostics for many chains: split-R-hat and bulk ESS.

Per-chain moments and autocovariance-based ESS are computed on the device by
``chain_diag_kernel`` (librsfm); only O(d) sums cross ranks, through one
all-reduce each.  The pure arithmetic that turns pooled
sums into R-hat lives in ``rhat_from_sums`` so that it can be tested on CPU.
"""
import numpy as np

from . import _lib
from .sharding import all_reduce_sum_


def rhat_from_sums(m, n, sum_mean, sum_mean2, sum_var):
    """Split-R-hat from pooled sums over m half-chains of length n.

    W = mean within-chain variance, B/n = variance of the half-chain means;
    R-hat = sqrt(((n-1)/n * W + B/n) / W)   (Gelman et al., BDA3 section 11.4).
    """
    w = sum_var / m
    var_means = (sum_mean2 - sum_mean * sum_mean / m) / (m - 1.0)
    if not w > 0:
        return float("nan")
    return float(np.sqrt(((n - 1.0) / n * w + var_means) / w))


def chain_diagnostics(samples, max_lag=None):
    """samples: CUDA tensor [n, d, C_local] (post-burn-in).  Returns a dict with per-parameter
    ``rhat``, ``ess`` (sum over all chains of all ranks), ``ess_per_chain_mean``, ``mean``, ``sd``."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n, d, c = samples.shape
    dev = samples.device
    samples = samples.contiguous()
    half = n // 2
    if max_lag is None:
        max_lag = min(n - 1, 250)
    out = {"n": int(n), "n_chains_local": int(c), "rhat": [], "ess": [], "ess_per_chain_mean": [],
           "mean": [], "sd": []}
    with torch.cuda.device(dev):
        stream = _lib.current_stream(torch, dev)
        mean = torch.empty(c, dtype=torch.float64, device=dev)
        var = torch.empty(c, dtype=torch.float64, device=dev)
        ess = torch.empty(c, dtype=torch.float64, device=dev)
        for p in range(d):
            sums = torch.zeros(8, dtype=torch.float64, device=dev)
            for h, (lo, hi) in enumerate(((0, half), (n - half, n))):
                if half < 2:
                    break                                   # too few draws for split-R-hat (reported as NaN)
                _lib.check(lib.rsfm_chain_diagnostics(_lib.ptr(samples[lo:hi]), hi - lo, d, c, p, 0,
                                                      _lib.ptr(mean), _lib.ptr(var), None, stream),
                           "rsfm_chain_diagnostics")
                sums[0] += c
                sums[1] += mean.sum()
                sums[2] += (mean * mean).sum()
                sums[3] += var.sum()
            _lib.check(lib.rsfm_chain_diagnostics(_lib.ptr(samples), n, d, c, p, int(max_lag), _lib.ptr(mean),
                                                  _lib.ptr(var), _lib.ptr(ess), stream), "rsfm_chain_diagnostics")
            sums[4] = ess.sum()
            sums[5] = c
            sums[6] = mean.sum()
            sums[7] = (var * (n - 1) + n * mean * mean).sum()        # sum of x^2 over all draws
            all_reduce_sum_(sums)
            s = sums.cpu().numpy()
            out["rhat"].append(rhat_from_sums(s[0], half, s[1], s[2], s[3]) if half >= 2 else float("nan"))
            out["ess"].append(float(s[4]))
            out["ess_per_chain_mean"].append(float(s[4] / s[5]))
            gmean = s[6] / s[5]
            out["mean"].append(float(gmean))
            out["sd"].append(float(np.sqrt(max(s[7] / (s[5] * n) - gmean * gmean, 0.0))))
    return out
