"""One rsf_forward_kernel launch in the stiff regime (cfg-4 style, reduced) for profiling."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
n = int(os.environ.get("STIFF_N", "4000"))
m = pkg.RateStateModel(number_time_steps=n, end_time=n * 0.1)
m.loading, m.vstep_period, m.vstep_factor = "vstep", 100.0, 10.0
dc = torch.from_numpy(np.random.default_rng(0).uniform(0.03, 0.08, int(os.environ.get("STIFF_C", "2048")))).cuda()
for _ in range(2):
    out = m.evaluate_batch(dc, want_acc=False, data=np.zeros(m.num_outputs()))
torch.cuda.synchronize()
print(float(out["sse"].sum()), int(out["nrhs"].sum()), int((out["status"] != 0).sum()))
