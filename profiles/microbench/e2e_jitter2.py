"""Which host-side call absorbs the spread of the public-API call?  cProfile of the slow calls only."""
import cProfile, importlib, io, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
m = pkg.RateStateModel(); m.Dc = 1325.0
np.random.seed(2024)
_, _, data = m.evaluate()
q0 = np.random.default_rng(1).uniform(200, 5000, 1024); q0[0] = 1000.0
def call():
    mc = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0, nsamples=2000, n_chains=1024, verbose=False, seed=1)
    return mc.sample(False)
call()
shown = 0
for i in range(24):
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable(); call(); pr.disable(); w = time.perf_counter() - t0
    print(f"call {i} wall {w*1e3:.1f} ms")
    if w > 0.45 and shown < 4:
        shown += 1
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(6); print(s.getvalue()[-1400:])
