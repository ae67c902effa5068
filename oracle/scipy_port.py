"""NumPy/SciPy restatement of the reference hot path (CPU ORACLE, test infrastructure).

This is the second form of the oracle: where rsf_oracle.c restates SciPy's DOP853
itself, this module drives the SAME third-party solver the reference drives --
``scipy.integrate.ode('dop853', rtol=1e-6, atol=1e-10)``, one ``integrate`` call
per output interval (RateStateModel.py:374-389) -- through a Python right-hand
side built from NumPy scalar ufuncs, as the reference's ``from numpy import exp,
log, sin`` does (imports.py:6, quirk q13).  It therefore has the reference's CPU
performance characteristics (about 90 % of the time in the Python callback,
SURVEY.md 3.3) and is what ``bench.py`` times as the CPU baseline ("port") and as
``--impl reference`` on the GPU box, where /root/reference does not exist.

Pinned by tests/test_oracle_golden.py::test_scipy_port_* against the golden
vectors generated from the unmodified reference.

Allowed importers: tests/, bench.py (cpu_baseline and --impl reference legs).
"""
import numpy as np
from numpy import exp, log, sin
from scipy import integrate


class PortModel:
    """Constants of RateStateModel.py:5-11,167-184."""

    def __init__(self, number_time_steps=500, start_time=0.0, end_time=50.0):
        self.a, self.b, self.mu_ref, self.V_ref, self.k1 = 0.011, 0.014, 0.6, 1.0, 1.0e-7
        self.t_start, self.t_final, self.num_tsteps = start_time, end_time, number_time_steps
        self.delta_t = (end_time - start_time) / number_time_steps
        self.mu_t_zero = 0.6
        self.RadiationDamping = True
        self.Dc = None
        self.n_rhs = 0
        # extensions of the build; the defaults are the reference
        self.loading = "sine_decay"          # "vstep": ONLY the load-point line of the RHS is swapped
        self.vstep_period, self.vstep_factor = 1000.0, 10.0
        self.observable = "acc"              # "mu": evaluate()[1] is the friction series (:385)
        self.state_law = "aging"             # "slip": ONLY the state-evolution line (:340) is swapped (Ruina)
        self.load_table, self.load_dt = None, 0.1   # loading = "table": piecewise-linear V_l/V_ref - 1

    def _rhs(self, t, y):
        # RateStateModel.py:318-355
        self.n_rhs += 1
        a, b, dc, v_ref = self.a, self.b, self.Dc, self.V_ref
        kprime = 1e-2 * 10 / dc
        if self.loading == "vstep":
            odd = int(np.floor((t - self.t_start) / self.vstep_period)) & 1
            v_l = self.vstep_factor * v_ref if odd else v_ref
        elif self.loading == "table":
            tab = self.load_table
            x = (t - self.t_start) / self.load_dt
            fi = min(max(np.floor(x), 0.0), float(len(tab) - 2))
            fr = min(max(x - fi, 0.0), 1.0)
            i = int(fi)
            v_l = v_ref * (1 + (tab[i] + fr * (tab[i + 1] - tab[i])))
        else:
            v_l = v_ref * (1 + exp(-t / 20) * sin(10 * t))            # RateStateModel.py:327-329
        out = np.zeros((len(y), 1))
        temp = 1 / a * (y[0] - self.mu_ref - b * log(v_ref * y[1] / dc))
        v = v_ref * exp(temp)
        out[1] = 1. - v * y[1] / dc
        if self.state_law == "slip":
            z = v * y[1] / dc
            out[1] = -z * log(z)
        out[0] = kprime * v_l - kprime * v
        out[2] = v / a * (out[0] - b / y[1] * out[1])
        if self.RadiationDamping:
            out[0] = out[0] - self.k1 * out[2]
            out[2] = v / a * (out[0] - b / y[1] * out[1])
        return out

    def evaluate(self):
        """(t, acc, acc_noise) as RateStateModel.evaluate(), :357-395."""
        n = int(np.floor((self.t_final - self.t_start) / self.delta_t))
        t, vel, acc, mu = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
        t[0] = self.t_start
        vel[0] = self.V_ref
        mu[0] = self.mu_ref
        solver = integrate.ode(self._rhs).set_integrator("dop853", rtol=1e-6, atol=1e-10)
        # the reference stores Dc/V_ref into a float array slot first (:369), so y0 holds scalars even
        # when Dc is the 1-element array MCMC.SSqcalc sets (q6); the RHS keeps using the array
        theta0 = float(np.ravel(self.Dc / self.V_ref)[0])
        solver.set_initial_value([self.mu_t_zero, theta0, self.V_ref], t[0])
        k = 1
        while solver.successful() and k < n:
            solver.integrate(solver.t + self.delta_t)
            t[k] = solver.t
            mu[k] = solver.y[0]
            vel[k] = solver.y[2]
            acc[k] = (vel[k] - vel[k - 1]) / self.delta_t
            k += 1
        acc_noise = acc + 1.0 * np.abs(acc) * np.random.randn(acc.shape[0])
        if self.observable == "mu":
            return t, mu, mu + np.abs(mu - mu[0]) * np.random.randn(n)
        return t, acc, acc_noise


def sse(model, q, data):
    """MCMC.SSqcalc (MCMC.py:381-389); Dc is set to a 1-element array like the reference (q6)."""
    model.Dc = np.asarray(q, dtype=np.float64).reshape(1)
    acc = model.evaluate()[1]
    return float(np.sum((acc.reshape(1, -1) - data) ** 2, axis=1)[0])


def run_chain(data, qstart, lo, hi, nsamples, n_prior_len=3, seed=None, model=None):
    """MCMC.sample(False) for list-typed priors (no adaptation, q2); SURVEY.md Appendix A.

    Returns dict(chain [nsamples+1], std2 [nsamples+1], accepts, n_solves, n_rhs)."""
    import warnings
    from scipy.stats import gamma
    if seed is not None:
        np.random.seed(seed)
    model = model or PortModel(number_time_steps=len(data))
    n0, n = 0.01, len(data)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.Dc = qstart
        acc = model.evaluate()[1]
        model.Dc *= (1 + 1e-6)
        acc_dq = model.evaluate()[1]
        std2 = [float(np.sum((acc - data) ** 2)) / (n - n_prior_len)]
        x = (acc_dq - acc) / (model.Dc * 1e-6)
        v = std2[0] / float(x @ x)
        ss = sse(model, qstart, data)
        solves = 3
        chain = [float(qstart)]
        accepts = []
        for _ in range(nsamples):
            qn = float(np.random.multivariate_normal([chain[-1]], [[v]])[0])
            ok = lo < qn < hi
            if ok:
                ssn = sse(model, qn, data)
                solves += 1
                ok = min(0.0, 0.5 * (ss - ssn) / std2[-1]) > np.log(np.random.rand(1))[0]
                if ok:
                    ss = ssn
            chain.append(qn if ok else chain[-1])
            accepts.append(bool(ok))
            std2.append(1 / gamma.rvs(0.5 * (n0 + n), scale=1 / (0.5 * (n0 * std2[-1] + ss)), size=1)[0])
    return {"chain": np.array(chain), "std2": np.array(std2), "accepts": np.array(accepts),
            "n_solves": solves, "n_rhs": model.n_rhs}


def run_chain_abdc(data, qstart, lo, hi, nsamples, step_sd, seed=None, model=None):
    """The same loop for the joint (a, b, Dc) posterior (extension, SURVEY 8f.4 / cfg 3): q = (a, b, Dc), per-parameter
    strict bounds, a fixed diagonal random-walk proposal (a bounded CPU sample cannot pool a covariance over
    65,536 chains; the cost per iteration -- one forward solve per in-bounds proposal -- is what is timed)."""
    import warnings
    from scipy.stats import gamma
    if seed is not None:
        np.random.seed(seed)
    model = model or PortModel(number_time_steps=len(data))
    n0, n = 0.01, len(data)
    lo, hi, step_sd = (np.asarray(x, dtype=np.float64) for x in (lo, hi, step_sd))

    def sse3(q):
        model.a, model.b = float(q[0]), float(q[1])
        return sse(model, q[2], data)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        q = np.asarray(qstart, dtype=np.float64).copy()
        ss = sse3(q)
        std2 = [ss / (n - 3)]
        solves = 1
        chain, accepts = [q.copy()], []
        for _ in range(nsamples):
            qn = q + step_sd * np.random.randn(3)
            ok = bool(np.all((qn > lo) & (qn < hi)))
            if ok:
                ssn = sse3(qn)
                solves += 1
                ok = min(0.0, 0.5 * (ss - ssn) / std2[-1]) > np.log(np.random.rand(1))[0]
                if ok:
                    q, ss = qn, ssn
            chain.append(q.copy())
            accepts.append(ok)
            std2.append(1 / gamma.rvs(0.5 * (n0 + n), scale=1 / (0.5 * (n0 * std2[-1] + ss)), size=1)[0])
    return {"chain": np.array(chain), "std2": np.array(std2), "accepts": np.array(accepts),
            "n_solves": solves, "n_rhs": model.n_rhs}


def _worker(args):
    if len(args) == 7:
        data, qstart, lo, hi, nsamples, seed, step_sd = args
        import time
        t0 = time.perf_counter()
        r = run_chain_abdc(data, qstart, lo, hi, nsamples, step_sd, seed=seed)
        return r["n_solves"], time.perf_counter() - t0, r["chain"]
    data, qstart, lo, hi, nsamples, seed = args
    import time
    t0 = time.perf_counter()
    r = run_chain(data, qstart, lo, hi, nsamples, seed=seed)
    return r["n_solves"], time.perf_counter() - t0, r["chain"]


def run_chains_parallel(data, qstarts, lo, hi, nsamples, seeds, processes, step_sd=None):
    """Independent chains on `processes` host cores (the reference itself is single-threaded and not
    re-entrant -- global RNG, shared model.Dc -- so separate processes are the only valid way).
    step_sd given: joint (a, b, Dc) chains (qstarts [C, 3], lo / hi [3])."""
    import multiprocessing as mp
    import time
    ctx = mp.get_context("fork")
    if step_sd is not None:
        jobs = [(data, np.asarray(q, dtype=np.float64), lo, hi, nsamples, int(s), step_sd) for q, s in zip(qstarts, seeds)]
    else:
        jobs = [(data, float(q), lo, hi, nsamples, int(s)) for q, s in zip(qstarts, seeds)]
    t0 = time.perf_counter()
    with ctx.Pool(processes) as pool:
        res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    return {"n_solves": sum(r[0] for r in res), "wall_s": wall, "chains": [r[2] for r in res]}
