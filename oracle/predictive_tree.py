"""CPU model of the speculative kernel's scheduling (test infrastructure; the product never imports it).

``rsf_mcmc_spec_kernel`` (csrc/rsfm_kernels.cu) evaluates G nodes of the tree of a chain's possible
futures per round, chosen best first by the probability that the realised path reaches them; the probabilities come
from a quadratic least-squares fit of the sum of squares in 1/Dc through every completed solve.  This module replays
that policy in plain Python on top of the C oracle's forward solve (oracle/rsf_oracle.c), so that two things can be
checked without a GPU:

  * the chain it produces is the one the sequential loop of MCMC.py:494-521 produces from the same draws, whatever
    the tree (the schedule never enters a result);
  * how many iterations a round advances with and without the fit (the design numbers quoted in DESIGN.md).

It mirrors the kernel's policy, not its arithmetic: sums are NumPy sums, the fit is ``numpy.linalg.solve``.
Allowed importers: tests/.
"""
import math

import numpy as np

from . import oracle as orc

N0 = 0.01          # MCMC.py:97


class _SS:
    """SS(q) = sum((model(q) - data)^2) through the C oracle, memoised (the tree re-visits proposals)."""

    def __init__(self, model, data, nthreads=4):
        self.model, self.data, self.nthreads, self.cache, self.solves = model, np.ascontiguousarray(data), nthreads, {}, 0

    def many(self, qs):
        need = [float(q) for q in qs if float(q) not in self.cache]
        if need:
            s = orc.forward_batch(self.model, np.array(need), data=self.data, nthreads=self.nthreads)[0]
            for q, v in zip(need, s):
                self.cache[q] = float(v)
            self.solves += len(need)
        return [self.cache[float(q)] for q in qs]


def sequential_chain(model, data, q0, var, lo, hi, z, u, gam):
    """The reference loop with absolute draws: proposals q + sqrt(var) z, strict box, log-ratio rule, sigma^2 draw
    (MCMC.py:497, 318-331, 158-160).  Returns (chain [n+1], accepts [n])."""
    ss_of = _SS(model, data)
    n = len(z)
    q, ss = float(q0), ss_of.many([q0])[0]
    s2 = ss / (len(data) - 3)
    chain, acc = [q], []
    for i in range(n):
        qn = q + math.sqrt(var) * z[i]
        ok = lo < qn < hi
        if ok:
            ssn = ss_of.many([qn])[0]
            ok = min(0.0, 0.5 * (ss - ssn) / s2) > math.log(u[i])
            if ok:
                q, ss = qn, ssn
        acc.append(bool(ok))
        chain.append(q)
        s2 = 1.0 / (gam[i] * (1.0 / (0.5 * (N0 * s2 + ss))))
    return np.array(chain), np.array(acc)


def speculative_chain(model, data, q0, var, lo, hi, z, u, gam, lanes, use_fit=True, forget=0.7):
    """The same chain by rounds of ``lanes`` concurrently evaluated tree nodes.  Returns (chain, accepts, stats) with
    stats = rounds, mean advance per round, nodes evaluated, fraction of decisions the policy predicted."""
    ss_of = _SS(model, data)
    n, sd = len(z), math.sqrt(var)
    q, ss = float(q0), ss_of.many([q0])[0]
    s2 = ss / (len(data) - 3)
    qc, xs, ss0 = float(q0), float(q0) * float(q0) / sd, ss
    mom, coef, tau2 = np.zeros(8), None, None
    chain, acc = [q], []
    it = rounds = evaluated = predicted = decided = 0
    while it < n:
        rmax = min(lanes, n - it)
        fit = coef is not None and tau2 is not None and use_fit
        p_prior = min(max((sum(acc) + 1.0) / (decided + 2.0), 0.05), 0.95)

        def node(cur, ssc, s2c, depth, reach, exact):
            m = it + depth - 1
            qn = cur + sd * z[m]
            inb = lo < qn < hi
            thr = ssc - 2.0 * s2c * math.log(u[m])
            sshat, p = ssc, p_prior
            if not inb:
                p = 0.0
            elif fit and qn > 0:
                x = (1.0 / qn - 1.0 / qc) * xs
                sshat = ss0 + coef[0] + x * (coef[1] + x * coef[2])
                p = min(max(0.5 * math.erfc(-(thr - sshat) / math.sqrt(2.0 * max(tau2, 1e-300))), 0.02), 0.98)
            kids = depth < rmax
            return dict(cur=cur, ssc=ssc, s2c=s2c, depth=depth, reach=reach, exact=exact, qn=qn, inb=inb, p=p, sshat=sshat,
                        g=gam[m], vA=reach * p if kids else -1.0, vR=reach * (1.0 - p) if kids else -1.0, cA=-1, cR=-1)

        nodes = [node(q, ss, s2, 1, 1.0, True)]
        for _ in range(1, lanes):
            best, bi, bc = 0.0, -1, None
            for i, nd in enumerate(nodes):                       # ties: lowest slot, accept child first (as the kernel)
                for c in ("A", "R"):
                    if nd["v" + c] > best:
                        best, bi, bc = nd["v" + c], i, c
            if bi < 0:
                break
            par = nodes[bi]
            par["v" + bc], par["c" + bc] = -1.0, len(nodes)
            ssc = par["sshat"] if bc == "A" else par["ssc"]
            s2n = 1.0 / (par["g"] * (1.0 / (0.5 * (N0 * par["s2c"] + ssc))))
            nodes.append(node(par["qn"] if bc == "A" else par["cur"], ssc, s2n, par["depth"] + 1,
                              par["reach"] * (par["p"] if bc == "A" else 1.0 - par["p"]), par["exact"] and bc == "R"))
        vals = ss_of.many([nd["qn"] for nd in nodes if nd["inb"]])
        evaluated += len(vals)
        k = 0
        for nd in nodes:
            if nd["inb"]:
                nd["sse"], k = vals[k], k + 1
        cl = done = 0
        while cl >= 0 and done < rmax:                            # the realised path through the tree
            nd, m = nodes[cl], it + done
            ok = False
            if nd["inb"]:
                ok = min(0.0, 0.5 * (ss - nd["sse"]) / s2) > math.log(u[m])
                decided += 1
                predicted += int(ok == (nd["p"] > 0.5))
                if ok:
                    q, ss = nd["qn"], nd["sse"]
            acc.append(bool(ok))
            chain.append(q)
            s2 = 1.0 / (gam[m] * (1.0 / (0.5 * (N0 * s2 + ss))))
            done += 1
            cl = nd["cA"] if ok else nd["cR"]
        it += done
        rounds += 1
        pts = [((1.0 / nd["qn"] - 1.0 / qc) * xs, nd["sse"] - ss0) for nd in nodes if nd["inb"] and nd["qn"] > 0]
        if coef is not None and pts:
            msr = float(np.mean([(y - (coef[0] + x * (coef[1] + x * coef[2]))) ** 2 for x, y in pts]))
            tau2 = msr if tau2 is None else 0.7 * tau2 + 0.3 * msr
        mom = forget * mom + sum((np.array([1.0, x, x * x, x ** 3, x ** 4, y, x * y, x * x * y]) for x, y in pts), np.zeros(8))
        coef = None
        if mom[0] >= 6.0:
            a = np.array([[mom[0], mom[1], mom[2]], [mom[1], mom[2], mom[3]], [mom[2], mom[3], mom[4]]])
            try:
                c = np.linalg.solve(a, mom[5:8])
                coef = c if c[2] > 0 else None
            except np.linalg.LinAlgError:
                coef = None
        if coef is not None and tau2 is None and len(pts) > 3:     # first fit: in-sample residual, n - 3 degrees of freedom
            tau2 = float(sum((y - (coef[0] + x * (coef[1] + x * coef[2]))) ** 2 for x, y in pts)) / (len(pts) - 3.0)
    stats = dict(rounds=rounds, advance=n / rounds, evaluated=evaluated, predicted=predicted / max(1, decided))
    return np.array(chain), np.array(acc), stats


# ---------------------------------------------------------------------------------------------------------------
# joint (a, b, Dc) posterior: the same policy with a ten-coefficient quadratic in t = (1/a, b/a, 1/Dc)
# ---------------------------------------------------------------------------------------------------------------
class _SS3:
    def __init__(self, data):
        self.data, self.cache = np.ascontiguousarray(data), {}

    def many(self, qs):
        out = []
        for q in qs:
            k = tuple(float(x) for x in q)
            if k not in self.cache:
                self.cache[k] = orc.sse(orc.forward(orc.make_model(Dc=k[2], a=k[0], b=k[1]))[1], self.data)
            out.append(self.cache[k])
        return out


def _features(q, qc, sc):
    t = np.array([1.0 / q[0] - 1.0 / qc[0], q[1] / q[0] - qc[1] / qc[0], 1.0 / q[2] - 1.0 / qc[2]]) * sc
    return np.array([1.0, t[0], t[1], t[2], t[0] * t[0], t[0] * t[1], t[0] * t[2], t[1] * t[1], t[1] * t[2], t[2] * t[2]])


def chain_abdc(data, q0, sd, lo, hi, z, u, gam, lanes=0, use_fit=True, forget=0.7):
    """Joint (a, b, Dc) chain with independent proposal s.d. ``sd`` [3] and draws z [n, 3], u [n], gam [n]:
    ``lanes = 0`` runs the sequential loop, otherwise rounds of ``lanes`` tree nodes.  Returns (chain [n+1, 3], accepts
    [n], stats)."""
    ss_of = _SS3(data)
    lo, hi, sd = np.asarray(lo, float), np.asarray(hi, float), np.asarray(sd, float)
    n = len(u)
    q = np.array(q0, dtype=float)
    ss = ss_of.many([q])[0]
    s2 = ss / (len(data) - 3)
    chain, acc = [q.copy()], []
    if lanes == 0:
        for i in range(n):
            qn = q + sd * z[i]
            ok = bool(np.all(qn > lo) and np.all(qn < hi))
            if ok:
                ssn = ss_of.many([qn])[0]
                ok = min(0.0, 0.5 * (ss - ssn) / s2) > math.log(u[i])
                if ok:
                    q, ss = qn, ssn
            acc.append(ok)
            chain.append(q.copy())
            s2 = 1.0 / (gam[i] * (1.0 / (0.5 * (N0 * s2 + ss))))
        return np.array(chain), np.array(acc), dict(rounds=n, advance=1.0, evaluated=n, predicted=0.0)
    qc, ss0 = q.copy(), ss
    sc = np.array([qc[0] ** 2 / sd[0], qc[0] / sd[1], qc[2] ** 2 / sd[2]])
    pm, rv, coef, tau2 = np.zeros((10, 10)), np.zeros(10), None, None
    it = rounds = evaluated = predicted = decided = 0
    while it < n:
        rmax = min(lanes, n - it)
        fit = use_fit and coef is not None and tau2 is not None
        p_prior = min(max((sum(acc) + 1.0) / (decided + 2.0), 0.05), 0.95)

        def node(cur, ssc, s2c, depth, reach):
            m = it + depth - 1
            qn = cur + sd * z[m]
            inb = bool(np.all(qn > lo) and np.all(qn < hi))
            thr = ssc - 2.0 * s2c * math.log(u[m])
            sshat, p = ssc, p_prior
            if not inb:
                p = 0.0
            elif fit:
                sshat = ss0 + float(_features(qn, qc, sc) @ coef)
                p = min(max(0.5 * math.erfc(-(thr - sshat) / math.sqrt(2.0 * max(tau2, 1e-300))), 0.02), 0.98)
            kids = depth < rmax
            return dict(cur=cur, ssc=ssc, s2c=s2c, depth=depth, reach=reach, qn=qn, inb=inb, p=p, sshat=sshat, g=gam[m],
                        vA=reach * p if kids else -1.0, vR=reach * (1.0 - p) if kids else -1.0, cA=-1, cR=-1)

        nodes = [node(q, ss, s2, 1, 1.0)]
        for _ in range(1, lanes):
            best, bi, bc = 0.0, -1, None
            for i, nd in enumerate(nodes):
                for c in ("A", "R"):
                    if nd["v" + c] > best:
                        best, bi, bc = nd["v" + c], i, c
            if bi < 0:
                break
            par = nodes[bi]
            par["v" + bc], par["c" + bc] = -1.0, len(nodes)
            ssc = par["sshat"] if bc == "A" else par["ssc"]
            s2n = 1.0 / (par["g"] * (1.0 / (0.5 * (N0 * par["s2c"] + ssc))))
            nodes.append(node(par["qn"] if bc == "A" else par["cur"], ssc, s2n, par["depth"] + 1,
                              par["reach"] * (par["p"] if bc == "A" else 1.0 - par["p"])))
        vals = ss_of.many([nd["qn"] for nd in nodes if nd["inb"]])
        evaluated += len(vals)
        k = 0
        for nd in nodes:
            if nd["inb"]:
                nd["sse"], k = vals[k], k + 1
        cl = done = 0
        while cl >= 0 and done < rmax:
            nd, m = nodes[cl], it + done
            ok = False
            if nd["inb"]:
                ok = min(0.0, 0.5 * (ss - nd["sse"]) / s2) > math.log(u[m])
                decided += 1
                predicted += int(ok == (nd["p"] > 0.5))
                if ok:
                    q, ss = nd["qn"], nd["sse"]
            acc.append(ok)
            chain.append(q.copy())
            s2 = 1.0 / (gam[m] * (1.0 / (0.5 * (N0 * s2 + ss))))
            done += 1
            cl = nd["cA"] if ok else nd["cR"]
        it += done
        rounds += 1
        pts = [(_features(nd["qn"], qc, sc), nd["sse"] - ss0) for nd in nodes if nd["inb"]]
        if coef is not None and pts:
            msr = float(np.mean([(y - f @ coef) ** 2 for f, y in pts]))
            tau2 = msr if tau2 is None else 0.7 * tau2 + 0.3 * msr
        pm = forget * pm + sum((np.outer(f, f) for f, _ in pts), np.zeros((10, 10)))
        rv = forget * rv + sum((f * y for f, y in pts), np.zeros(10))
        coef = None
        if pm[0, 0] >= 20.0:
            try:
                coef = np.linalg.solve(pm + 1e-10 * np.diag(np.diag(pm)), rv)
            except np.linalg.LinAlgError:
                coef = None
        if coef is not None and tau2 is None and len(pts) > 10:
            tau2 = float(sum((y - f @ coef) ** 2 for f, y in pts)) / (len(pts) - 10.0)
    stats = dict(rounds=rounds, advance=n / rounds, evaluated=evaluated, predicted=predicted / max(1, decided))
    return np.array(chain), np.array(acc), stats
