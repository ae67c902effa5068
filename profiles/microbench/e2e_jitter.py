"""Is the run-to-run spread of the public-API call in the kernel or on the host?  (cfg 2, 1024 x 2000)"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
lib = pkg._lib.load()
orig = lib.rsfm_run
ev = []
def timed_run(*a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = orig(*a); e1.record(); ev.append((e0, e1)); return rc
lib.rsfm_run = timed_run
m = pkg.RateStateModel(); m.Dc = 1325.0
np.random.seed(2024)
_, _, data = m.evaluate()
q0 = np.random.default_rng(1).uniform(200, 5000, 1024); q0[0] = 1000.0
gap = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
for i in range(24):
    if gap: time.sleep(gap)
    t0 = time.perf_counter()
    mc = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0, nsamples=2000, n_chains=1024, verbose=False, seed=1)
    mc.sample(False)
    w = time.perf_counter() - t0
    torch.cuda.synchronize()
    k = sum(a.elapsed_time(b) for a, b in ev) / 1e3; ev.clear()
    print(f"call {i:2d} wall {w*1e3:7.1f} ms  rsfm_run on device {k*1e3:7.1f} ms  host+copies {1e3*(w-k):6.1f} ms")
