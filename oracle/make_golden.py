#!/usr/bin/env python
"""Generate tests/golden/*.json from the UNMODIFIED reference (container-only).

The reference holds no tests, fixtures or golden vectors,
so parity is pinned on outputs of the reference itself, imported from
/root/reference through oracle/ref_shim.py and run here:

  G1  forward_trajectories.json   RateStateModel.evaluate()[1]  (RateStateModel.py:188)
        Dc in {1,10,100,1000,1350,5000,10000} x RadiationDamping in {True, False},
        plus a stiff case (Dc = 0.05) and the silent-failure case (Dc = 1e-4, q9).
  G2  chain_list_priors.json / chain_dict_priors.json   MCMC.sample(False) (MCMC.py:391)
        run with np.random.seed(...), with every random draw logged
        (proposal, uniform, unit-scale gamma) plus accept flags, chain and std2.
  G3  sse_grid.json   SSE(Dc) on a grid for the seeded data set (MCMC.py:387).
  G4  rhs_values.json   the nested RHS ``friction(t, y)`` (RateStateModel.py:277-355) at fixed states
        (``python oracle/make_golden.py rhs``).
  G5  forward_k1.json   trajectories and SSE(k1) with the model attribute ``k1`` varied
        (``python oracle/make_golden.py k1``).

Everything is written in the reference's own ``__ndarray__`` JSON wire format
(json_save_load.py:37-38).  Versions of numpy/scipy are stamped into each file:
the reference pins none, and SciPy's dop853 is part of the oracle's identity.

Usage: python oracle/make_golden.py         (about two minutes)
"""
import contextlib
import io
import json
import os
import platform
import sys
import warnings

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STAMP = {"numpy": np.__version__, "scipy": scipy.__version__, "python": platform.python_version(),
         "generator": "oracle/make_golden.py", "source": "unmodified reference via oracle/ref_shim.py"}


def enc(o):
    if isinstance(o, np.ndarray):
        return {"__ndarray__": True, "data": o.tolist(), "shape": list(o.shape)}
    if isinstance(o, (np.floating, np.integer, np.bool_)):
        return o.item()
    raise TypeError(type(o))


def dump(name, obj):
    obj = dict(obj)
    obj["_stamp"] = STAMP
    path = os.path.join(OUT, name)
    with open(path, "w") as f:
        json.dump(obj, f, default=enc)
    print("wrote", path, os.path.getsize(path), "bytes")


def ref_forward(rsm, dc, damping=True, n=500):
    m = rsm.RateStateModel(number_time_steps=n)
    m.Dc = float(dc)
    m.RadiationDamping = damping
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t, acc, _ = m.evaluate()
    return t, acc


def make_data(rsm, dc_true, seed, n=500):
    """cfg-1 data set: seed immediately before evaluate(); keep acc_noise."""
    m = rsm.RateStateModel(number_time_steps=n)
    m.Dc = float(dc_true)
    np.random.seed(seed)
    t, acc, acc_noise = m.evaluate()
    return acc_noise


def record_chain(rsm, mcm, data, dc_true, qpriors, qstart, nsamples, seed):
    """Run MCMC.sample(False) with all random draws logged."""
    log = {"proposals": [], "V_used": [], "uniforms": [], "gammas_unit": [], "gammas_scaled": []}
    real_mvn, real_rand, real_gamma = np.random.multivariate_normal, np.random.rand, mcm.gamma

    def mvn(mean, cov, *a, **k):
        x = real_mvn(mean, cov, *a, **k)
        log["proposals"].append(float(np.ravel(x)[0]))
        log["V_used"].append(float(np.ravel(cov)[0]))
        return x

    def rand(*a):
        u = real_rand(*a)
        log["uniforms"].append((len(log["proposals"]) - 1, float(np.ravel(u)[0])))
        return u

    class GammaProxy:
        """scipy.stats.gamma.rvs(a, scale=s) == s * RandomState.standard_gamma(a)."""
        @staticmethod
        def rvs(aval, scale=1.0, size=1):
            g0 = np.random.standard_gamma(aval, size)
            log["gammas_unit"].append(float(g0[0]))
            g = np.ravel(g0 * scale)          # scale arrives as a (1,1) array (SSq keeps dims)
            log["gammas_scaled"].append(float(g[0]))
            return g

    def run(patched):
        model = rsm.RateStateModel(number_time_steps=len(data))
        np.random.seed(seed)
        mc = mcm.MCMC(model, data, dc_true, qpriors, qstart, nsamples=nsamples)
        if patched:
            np.random.multivariate_normal, np.random.rand, mcm.gamma = mvn, rand, GammaProxy
        try:
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out = mc.sample(False)
        finally:
            np.random.multivariate_normal, np.random.rand, mcm.gamma = real_mvn, real_rand, real_gamma
        # MCMC.py:503 prints "<isample> <accept>"; accept is a (1,1) bool array when the
        # proposal was in bounds ("[[ True]]") and a numpy bool otherwise ("False")
        accepts = ["True" in ln for ln in buf.getvalue().splitlines()
                   if ln and ln.split()[0].isdigit() and ("True" in ln or "False" in ln)]
        assert len(accepts) == nsamples
        return out, np.asarray(mc.std2), np.asarray(mc.Vstart), accepts

    plain = run(False)
    rec = run(True)
    # the logging wrappers must not change the chain
    assert np.array_equal(plain[0], rec[0]) and np.array_equal(plain[1], rec[1]), "recorder changed the chain"
    out, std2, vstart, accepts = rec
    uni = np.full(nsamples, np.nan)
    for i, u in log["uniforms"]:
        uni[i] = u
    return {
        "seed": seed, "dc_true": dc_true, "qstart": qstart, "nsamples": nsamples,
        "nburn": int(nsamples / 2),
        "qpriors_form": "dict" if isinstance(qpriors, dict) else "list",
        "lo": float(qpriors[1]), "hi": float(qpriors[2]), "n_prior_len": len(qpriors),
        "data": np.asarray(data),
        "proposals": np.array(log["proposals"]), "V_used": np.array(log["V_used"]),
        "uniforms": uni, "gammas_unit": np.array(log["gammas_unit"]),
        "accepts": np.array(accepts, dtype=np.int64),
        "chain_post_burn": np.asarray(out), "std2_post_burn": std2, "Vstart": vstart,
    }


def capture_friction(rsm, model):
    """The reference's RHS is a function nested inside evaluate() (RateStateModel.py:277-355); it is handed to
    ``integrate.ode(friction)`` (:374).  Swapping that class for a recorder for the duration of ONE evaluate()
    call yields the unmodified closure, callable at any (t, y); the integration loop sees successful() False
    and exits at once."""
    got = {}

    class Recorder:
        def __init__(self, f):
            got["f"] = f

        def set_integrator(self, *a, **k):
            return self

        def set_initial_value(self, *a, **k):
            return self

        def successful(self):
            return False

    real = rsm.integrate.ode
    rsm.integrate.ode = Recorder
    try:
        model.evaluate()
    finally:
        rsm.integrate.ode = real
    return got["f"]


def make_rhs_values(rsm):
    """G4 rhs_values.json: friction(t, y) of the unmodified reference at fixed (t, y, Dc), both damping modes:
    states near sliding steady state (where the CUDA path uses its short series) and far from it (where it
    uses the general-range formulas)."""
    rng = np.random.default_rng(2718)
    rows = []
    for damping in (True, False):
        for dc in (0.05, 1.0, 50.0, 130.0, 1000.0, 1350.0, 10000.0):
            m = rsm.RateStateModel()
            m.RadiationDamping = damping
            m.Dc = dc
            f = capture_friction(rsm, m)
            for near in (True, False):
                for _ in range(6):
                    t = float(rng.uniform(0.0, 50.0))
                    if near:
                        th = dc * (1.0 + rng.uniform(-1.5e-4, 1.5e-4))
                        mu = 0.6 + rng.uniform(-3e-5, 3e-5)
                    else:
                        th = dc * float(np.exp(rng.uniform(-1.0, 1.0)))
                        mu = 0.6 + rng.uniform(-0.02, 0.02)
                    y = np.array([mu, th, 1.0 + rng.uniform(-1e-3, 1e-3)])
                    out = np.asarray(f(t, y), dtype=np.float64).reshape(3)
                    rows.append([float(damping), dc, t, y[0], y[1], y[2], out[0], out[1], out[2]])
    return {"columns": ["RadiationDamping", "Dc", "t", "mu", "theta", "V", "dmu", "dtheta", "dV"],
            "rows": np.array(rows)}


def make_k1_values(rsm):
    """G5 forward_k1.json: RateStateModel.evaluate()[1] of the unmodified reference with its public attribute ``k1``
    (RateStateModel.py:171, used at :351) moved away from the module default, and the sum of squares against a seeded
    data set on a k1 grid -- what pins the oracle (and through it the CUDA path) for chains that sample k1."""
    cases = []
    for dc in (100.0, 1000.0, 1350.0):
        for k1 in (0.0, 1e-7, 1e-5, 1e-3, 3e-3, 8e-3):
            m = rsm.RateStateModel(number_time_steps=500)
            m.Dc = dc
            m.k1 = k1
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                t, acc, _ = m.evaluate()
            cases.append({"Dc": dc, "k1": k1, "N": 500, "acc": acc, "t_last": float(t[-1])})
    # seeded data set at (Dc, k1) = (1000, 3e-3); SSE(k1) as MCMC.SSqcalc forms it (MCMC.py:387), model.k1 set instead of model.Dc
    m = rsm.RateStateModel(number_time_steps=500)
    m.Dc = 1000.0
    m.k1 = 3e-3
    np.random.seed(4242)
    _, _, data = m.evaluate()
    grid = np.array([1e-4, 5e-4, 1e-3, 2e-3, 2.5e-3, 3e-3, 3.5e-3, 4e-3, 6e-3, 9e-3])
    sse = []
    for k1 in grid:
        m.k1 = float(k1)
        acc = m.evaluate()[1]
        sse.append(float(np.sum((acc - data) ** 2, axis=0)))
    return {"cases": cases, "data": data, "data_Dc": 1000.0, "data_k1": 3e-3, "data_seed": 4242,
            "k1_grid": grid, "sse": np.array(sse)}


def main():
    os.makedirs(OUT, exist_ok=True)
    rsm, mcm = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "k1":
        dump("forward_k1.json", make_k1_values(rsm))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "rhs":
        dump("rhs_values.json", make_rhs_values(rsm))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "cfg1":
        data_cfg1 = make_data(rsm, 1325.0, 2024)
        dump("chain_cfg1_full.json",
             record_chain(rsm, mcm, data_cfg1, 1325.0, ["Uniform", 0.0, 10000.0], 1000.0, nsamples=500, seed=2024))
        return

    # G1
    g1 = {"cases": []}
    for damping in (True, False):
        for dc in (1.0, 10.0, 100.0, 1000.0, 1350.0, 5000.0, 10000.0):
            t, acc = ref_forward(rsm, dc, damping)
            g1["cases"].append({"Dc": dc, "RadiationDamping": damping, "N": 500, "acc": acc,
                                "t_last": float(t[-1])})
    t, acc = ref_forward(rsm, 0.05, True)
    g1["cases"].append({"Dc": 0.05, "RadiationDamping": True, "N": 500, "acc": acc, "t_last": float(t[-1])})
    t, acc = ref_forward(rsm, 1e-4, True)
    g1["cases"].append({"Dc": 1e-4, "RadiationDamping": True, "N": 500, "acc": acc, "t_last": float(t[-1]),
                        "filled": int(np.count_nonzero(t) + 1), "t1": float(t[1])})
    # a different grid: N = 200 over [0, 30]
    m = rsm.RateStateModel(number_time_steps=200, end_time=30.0)
    m.Dc = 700.0
    t, acc, _ = m.evaluate()
    g1["cases"].append({"Dc": 700.0, "RadiationDamping": True, "N": 200, "end_time": 30.0, "acc": acc,
                        "t_last": float(t[-1])})
    dump("forward_trajectories.json", g1)

    # G3 (and the data set used by G2)
    data = make_data(rsm, 1350.0, 12345)
    grid = np.array([300.0, 700.0, 1000.0, 1200.0, 1300.0, 1350.0, 1400.0, 1500.0, 2000.0, 4000.0, 9000.0])
    sse = []
    for dc in grid:
        _, acc = ref_forward(rsm, dc, True)
        sse.append(np.sum((acc.reshape(1, -1) - data) ** 2, axis=1).item())     # MCMC.py:387
    dump("sse_grid.json", {"dc_true": 1350.0, "seed": 12345, "data": data, "grid": grid, "sse": np.array(sse)})

    # G2
    dump("chain_list_priors.json",
         record_chain(rsm, mcm, data, 1350.0, ["Uniform", 0.0, 10000.0], 1000.0, nsamples=120, seed=2024))
    dump("chain_dict_priors.json",
         record_chain(rsm, mcm, data, 1350.0, {1: 0.0, 2: 10000.0}, 1000.0, nsamples=60, seed=7))
    # the reference's own default run in full (main.py:50-56, Dc_true = 1325): 500 iterations, data made with
    # np.random.seed(2024) immediately before evaluate() (SURVEY 8d cfg 1)
    data_cfg1 = make_data(rsm, 1325.0, 2024)
    dump("chain_cfg1_full.json",
         record_chain(rsm, mcm, data_cfg1, 1325.0, ["Uniform", 0.0, 10000.0], 1000.0, nsamples=500, seed=2024))
    # a chain that starts near the upper bound so that out-of-bounds proposals occur (q10)
    dump("chain_bounds.json",
         record_chain(rsm, mcm, data, 1350.0, ["Uniform", 900.0, 1500.0], 1450.0, nsamples=60, seed=99))
    dump("rhs_values.json", make_rhs_values(rsm))


if __name__ == "__main__":
    main()
