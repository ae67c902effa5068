"""Pooled adaptive-Metropolis proposal from sufficient statistics (extension).

The reference's own update (MCMC.py:162-204) is per chain, uses the last 10
samples and is either dead (list priors, q2) or dimensionally wrong (dict priors,
q3); it is reproduced on the device by ``RSFM_ADAPT_COMPAT``.  This module is the
corrected generalisation the north star asks for (Haario et al. 2001): the
covariance of ALL chains' draws, pooled over ranks by one all-reduce of
``[n, sum q (d), sum q q^T (lower triangle, row-major)]``, scaled by 2.38^2/d.
"""
import numpy as np

from . import _lib
from .sharding import all_gather_rows, max_pool_groups

EPS_REL = 1e-10


class PooledAdaptation:
    """Device-side pooled adaptive Metropolis around a ``rsfm_sampler`` (generalises
    MCMC.py:162-204, 523-527).  Per interval j of ``adapt_interval`` iterations, with no host synchronisation:

        main stream   [install(j-2)] [form(j-1)]  rsfm_run(j)  rsfm_pooled_partials(j)
        side stream                                            all-gather(j) ............

    ``partials`` are sums over fixed groups of 1,024 chains aligned on the global chain id; the all-gather
    (NCCL over NVLink; nothing for one rank) puts every rank's rows in global chain order and
    ``rsfm_pooled_update`` adds them to the running moments in that order, forms (2.38^2/d) cov and its
    Cholesky factor in closed form and installs it for every chain -- the same bits on every rank and for any
    number of ranks.  The all-gather of interval j overlaps the kernel of interval j+1; its factor is FORMED before
    interval j+1 is launched and used from interval j+2 on (adaptation lags one interval, as SURVEY 8e allows).
    Forming early matters with chain groups (``rsfm_chain_groups``: the chains of a large sampler are served by
    several launches on the sampler's own streams): a group's install waits for the forming of the factor, which
    happened an interval ago, and never for the other groups' launches still running.

    Use: ``before_interval()`` -> launch the interval's iterations on the current stream ->
    ``after_interval(end_iteration)``; ``finish()`` at the end."""

    def __init__(self, torch, lib, handle, dev, d, n_chains_total, world, adapt_start, stream):
        self.torch, self.lib, self.handle, self.dev, self.d = torch, lib, handle, dev, d
        self.world, self.adapt_start, self.stream = world, int(adapt_start), stream
        self.tri = d * (d + 1) // 2
        g, self.rows = _lib.POOL_GROUP, _lib.POOL_ROWS
        ng_local = int(lib.rsfm_pooled_groups(handle))
        ng = max_pool_groups(n_chains_total, world, g) if world > 1 else ng_local
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
        self.loc = [z(ng, self.rows), z(ng, self.rows)]
        self.parts = [z(world * ng, self.rows), z(world * ng, self.rows)] if world > 1 else self.loc
        self.moments = z(1 + d + self.tri)
        self.hist = z(64, 1 + self.tri)
        self.side = torch.cuda.Stream(dev) if world > 1 else None
        self.main = torch.cuda.current_stream(dev)
        self.gathered, self.ends, self.tev = {}, {}, []
        self.j = 0
        self.formed = -10                      # last interval whose factor has been formed

    def preload(self, moments, pending_rows, end_iteration):
        """State of a checkpoint taken on an adaptation boundary: the moments already applied and the gathered
        rows of the interval that ended there (they are applied one interval late)."""
        torch = self.torch
        pend = np.asarray(pending_rows, dtype=np.float64).reshape(-1, self.rows)
        if pend.shape[0] != self.parts[1].shape[0]:
            raise ValueError("checkpoint was written with a different number of ranks / chains")
        self.moments.copy_(torch.as_tensor(np.asarray(moments, dtype=np.float64)))
        self.parts[1].copy_(torch.as_tensor(pend))
        self.gathered[-1], self.ends[-1] = self.main.record_event(), int(end_iteration)

    def _form(self, j):
        """moments += rows(j) and, past adapt_start, form the factor they give (main stream); it is installed one
        interval later (``_install``)."""
        torch = self.torch
        self.main.wait_event(self.gathered[j])
        acc = 1 if self.ends[j] > self.adapt_start // 2 else 0          # the earliest draws stay out of the moments
        inst = 2 if self.ends[j] >= self.adapt_start else 0
        if j + 1 >= self.hist.shape[0]:
            self.hist = torch.cat([self.hist, torch.zeros_like(self.hist)])
        p = self.parts[j % 2]
        _lib.check(self.lib.rsfm_pooled_update(self.handle, _lib.ptr(p), int(p.shape[0]), _lib.ptr(self.moments), acc, inst,
                                               _lib.ptr(self.hist[j + 1]), self.stream), "rsfm_pooled_update")
        self.formed = j

    def _install(self):
        """The factor formed last (if any) replaces every chain's proposal factor."""
        _lib.check(self.lib.rsfm_pooled_install(self.handle, self.stream), "rsfm_pooled_install")

    def before_interval(self):
        j = self.j
        if self.formed == j - 2:
            self._install()                       # factor of interval j-2, formed before interval j-1 was launched
        if (j - 1) in self.gathered and self.formed < j - 1:
            self._form(j - 1)

    def after_interval(self, end_iteration):
        torch, j = self.torch, self.j
        self.ends[j] = int(end_iteration)
        _lib.check(self.lib.rsfm_pooled_partials(self.handle, _lib.ptr(self.loc[j % 2]), 1, self.stream), "rsfm_pooled_partials")
        ready = self.main.record_event()
        if self.world > 1:
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(self.side)
                all_gather_rows(self.parts[j % 2], self.loc[j % 2])
                t1.record(self.side)
                self.tev.append((t0, t1))
                self.gathered[j] = self.side.record_event()
        else:
            self.gathered[j] = ready
        self.j += 1

    def finish(self):
        """Apply what an uninterrupted run would have applied before its next interval; the last interval's rows
        stay pending (one-interval lag) and are returned with the moments for a checkpoint.  Synchronises."""
        torch = self.torch
        if self.formed == self.j - 2:
            self._install()
        pending = None
        if (self.j - 1) in self.gathered:
            self.main.wait_event(self.gathered[self.j - 1])
            pending = self.parts[(self.j - 1) % 2].clone()
        torch.cuda.synchronize(self.dev)
        h = self.hist.cpu().numpy()
        self.history = [(self.ends[k], h[k + 1, 1:].copy()) for k in sorted(self.ends) if h[k + 1, 0] == 1.0]
        self.stats = {"n_adaptations": len(self.history), "n_intervals": self.j,
                      "collective_ms_on_side_stream": float(sum(a.elapsed_time(b) for a, b in self.tev)),
                      "pool_rows_gathered": int(self.parts[0].shape[0])}
        return self.moments, pending


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
