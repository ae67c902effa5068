import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
c, iters = 131072, 400
m = pkg.RateStateModel(); m.Dc = 1325.0
np.random.seed(2024)
_, _, data = m.evaluate()
q0 = np.random.default_rng(0).uniform(200.0, 5000.0, c); q0[0] = 1000.0
for rep in range(3):
    mc = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0, nsamples=iters, n_chains=c, verbose=False, seed=5,
                  adapt="pooled", adapt_start=50)
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = mc.sample(False); w = time.perf_counter() - t0
    print(f"rep {rep}: wall {w:.3f} s, device loop (create..final sync) {mc.stats['elapsed_s']:.3f} s, results to host {w - mc.stats['elapsed_s']:.3f} s, "
          f"solves/s {mc.stats['nsolves']/w:.3e}")
    del out
