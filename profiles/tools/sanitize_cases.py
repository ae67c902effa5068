#!/usr/bin/env python
"""Small invocations of every kernel family, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python profiles/tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python profiles/tools/sanitize_cases.py
    compute-sanitizer --tool synccheck python profiles/tools/sanitize_cases.py

Covers the mbarrier / cp.async.bulk staging of the observed series in its three regimes (one resident tile: N = 500;
two resident tiles: N = 1,000; streamed double buffer with block barriers: N = 2,500), the cp.async prefetch ring of
the nominal load table, the speculative kernel's shuffles, the d = 3 out-of-bounds walk, the stiff variant's one-warp
blocks and the pooled-adaptation kernels.  Sizes are tiny: the tools slow the kernels down by 10-100x."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
import torch  # noqa: E402

rng = np.random.default_rng(0)


def data_for(model):
    np.random.seed(1)
    return model.evaluate()[2]


def forward(n, t_end, c, **attrs):
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    for k, v in attrs.items():
        setattr(m, k, v)
    m.Dc = 0.3 if "loading" in attrs else 1325.0
    d = data_for(m)
    o = m.evaluate_batch(rng.uniform(300.0, 4000.0, c) if "loading" not in attrs else rng.uniform(0.05, 2.0, c), data=d)
    torch.cuda.synchronize()
    assert np.all(o["status"].cpu().numpy() == 0)
    return m, d


def chains(m, d, c, ns, **kw):
    names = kw.get("param_names", ("Dc",))
    if len(names) == 3:
        q0 = np.stack([rng.uniform(0.0105, 0.0115, c), rng.uniform(0.0135, 0.0145, c), rng.uniform(900.0, 2000.0, c)], axis=1)
    elif getattr(m, "loading", "") == "vstep":
        q0 = rng.uniform(0.05, 0.4, c)
    else:
        q0 = rng.uniform(900.0, 2000.0, c)
    pri = kw.pop("qpriors", ["Uniform", 0.0, 1e4])
    mc = pkg.MCMC(m, d, 1325.0, pri, q0, nsamples=ns, n_chains=c, verbose=False, seed=3, **kw)
    out = mc.sample(False)
    assert np.all(np.isfinite(out))
    mc.diagnostics()
    return mc


print("forward N=500 / 1000 / 2500 (streamed)", flush=True)
m500, d500 = forward(500, 50.0, 70)
m1000, d1000 = forward(1000, 100.0, 70)
m2500, d2500 = forward(2500, 250.0, 70)
print("forward stiff variant, velocity steps", flush=True)
mv, dv = forward(1500, 150.0, 40, loading="vstep", vstep_period=30.0, vstep_factor=10.0)
print("mcmc sequential d=1: resident, two tiles, streamed", flush=True)
chains(m500, d500, 200, 6, spec_depth=1)
chains(m1000, d1000, 100, 4, spec_depth=1)
chains(m2500, d2500, 70, 4, spec_depth=1)
print("mcmc speculative d=1 (depth auto), one chain, compat adaptation", flush=True)
chains(m500, d500, 8, 12)
chains(m1000, d1000, 1, 8)
chains(m500, d500, 4, 24, qpriors={1: 0.0, 2: 1e4})
print("mcmc d=3: out-of-bounds walk and speculative, pooled adaptation", flush=True)
b3 = [[0.0100, 0.0120], [0.0130, 0.0150], [800.0, 2200.0]]
chains(m500, d500, 2048, 30, param_names=("a", "b", "Dc"), bounds=b3, adapt="pooled", adapt_start=10, spec_depth=1)
chains(m500, d500, 6, 8, param_names=("a", "b", "Dc"), bounds=b3)
print("mcmc stiff variant (one-warp blocks), streamed series", flush=True)
mv.Dc = 0.2
chains(mv, dv, 40, 3, qpriors=["Uniform", 0.01, 1.0], spec_depth=1)
print("kde", flush=True)
x = torch.from_numpy(rng.normal(1300.0, 60.0, 5000)).cuda()
pkg.gaussian_kde_pdf(x, np.linspace(1000.0, 1600.0, 1000))
torch.cuda.synchronize()
print("sanitize_cases: all cases ran", flush=True)
