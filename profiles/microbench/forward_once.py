"""One rsf_forward_kernel launch (SSE only) for profiling: C chains, default model."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
c = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = pkg.RateStateModel(); m.Dc = 1325.0
_, acc, _ = m.evaluate()
dc = torch.from_numpy(np.random.default_rng(0).uniform(800, 2000, c)).cuda()
for _ in range(3):
    out = m.evaluate_batch(dc, data=acc, want_acc=False)
torch.cuda.synchronize()
print(float(out["sse"].sum()))
