"""Stiff forward batch (cfg-4 style, reduced): how much of the lost SIMT efficiency is systematic, i.e. chains of
different stiffness (Dc) sharing a warp?  Same 16,384 parameter sets in random order and sorted by Dc.
usage: python profiles/microbench/forward_stiff_ab.py"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
n = 4000
m = pkg.RateStateModel(number_time_steps=n, end_time=n * 0.1)
m.loading, m.vstep_period, m.vstep_factor = "vstep", 100.0, 10.0
dc0 = np.random.default_rng(0).uniform(0.03, 0.08, 16384)
data = np.zeros(m.num_outputs())
for name, dc_h in (("random order", dc0), ("sorted by Dc", np.sort(dc0)), ("all Dc = 0.05", np.full(16384, 0.05))):
    dc = torch.from_numpy(dc_h).cuda()
    out = m.evaluate_batch(dc, want_acc=False, data=data)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = m.evaluate_batch(dc, want_acc=False, data=data)
    torch.cuda.synchronize()
    w = time.perf_counter() - t0
    print(f"{name:14s} {w*1e3:7.1f} ms  solves/s {dc.numel()/w:7.0f}  rhs {int(out['nrhs'].sum())}  steps {int(out['nstep'].sum())}"
          f"  rhs/s {int(out['nrhs'].sum())/w:.3e}  failed {int((out['status'] != 0).sum())}")
