"""Single-chain latency (cfg-1 style: one chain, 500 samples) with list priors (no adaptation) and dict priors (the
reference's adaptation every 10 samples), sequential kernel vs speculation."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
m = pkg.RateStateModel(); m.Dc = 1325.0
np.random.seed(2024)
_, _, data = m.evaluate()
for name, pri in (("list priors", ["Uniform", 0.0, 1e4]), ("dict priors", {1: 0.0, 2: 1e4})):
    for depth in (1, 0):
        ts = []
        for rep in range(4):
            mc = pkg.MCMC(m, data, 1325.0, pri, 1000.0, nsamples=500, verbose=False, seed=3, spec_depth=depth)
            t0 = time.perf_counter(); out = mc.sample(False); ts.append(time.perf_counter() - t0)
        print(f"{name}, spec_depth={depth}: {min(ts)*1e3:7.1f} ms for 500 samples  ({min(ts)/500*1e6:6.1f} us/iteration), "
              f"accept {mc.acceptance_ratio[0]:.2f}, mean {out.mean():.1f}")
