"""Wall time of every C-ABI call inside one public-API call (cfg 2), for the slow calls."""
import collections, importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
lib = pkg._lib.load()
acc = collections.OrderedDict()
def wrap(name):
    f = getattr(lib, name)
    def g(*a):
        t0 = time.perf_counter(); r = f(*a); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(lib, name, g)
for n in ("rsfm_create", "rsfm_init", "rsfm_run", "rsfm_get_totals", "rsfm_get_state", "rsfm_destroy", "rsfm_iteration"):
    wrap(n)
sync0 = torch.cuda.synchronize
def sync(*a, **k):
    t0 = time.perf_counter(); sync0(*a, **k); acc["synchronize"] = acc.get("synchronize", 0.0) + time.perf_counter() - t0
torch.cuda.synchronize = sync
m = pkg.RateStateModel(); m.Dc = 1325.0
np.random.seed(2024)
_, _, data = m.evaluate()
q0 = np.random.default_rng(1).uniform(200, 5000, 1024); q0[0] = 1000.0
for i in range(24):
    acc.clear(); t0 = time.perf_counter()
    mc = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0, nsamples=2000, n_chains=1024, verbose=False, seed=1)
    mc.sample(False); w = time.perf_counter() - t0
    print(f"call {i:2d} wall {w*1e3:7.1f} ms  " + "  ".join(f"{k[5:] if k.startswith('rsfm_') else k} {v*1e3:.1f}" for k, v in acc.items()))
