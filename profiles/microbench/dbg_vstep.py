import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
from oracle import oracle as orc
n = 20000
dcs = np.array([0.05, 0.1, 0.5, 1.0, 10.0, 100.0])
for period, factor in ((1000.0, 10.0), (1000.0, 1.0)):
    m = pkg.RateStateModel(number_time_steps=n, end_time=n * 0.1)
    m.loading, m.vstep_period, m.vstep_factor = "vstep", period, factor
    out = m.evaluate_batch(dcs, want_t=True)
    st, fl = out["status"].cpu().numpy(), out["filled"].cpu().numpy()
    acc = out["acc"].t().cpu().numpy(); tt = out["t"].t().cpu().numpy()
    print("period", period, "factor", factor, "status", st, "filled", fl, "nrhs", out["nrhs"].cpu().numpy())
    for i, dc in enumerate(dcs):
        om = orc.make_model(Dc=dc, number_time_steps=n, end_time=n * 0.1, loading=orc.LOAD_VSTEP, vstep_period=period, vstep_factor=factor)
        to, acco, sto = orc.forward(om)
        k = fl[i]
        scale = np.abs(acco).max()
        d = np.abs(acc[i, :k] - acco[:k]) / scale
        bad = np.nonzero(d > 1e-6)[0]
        print("  Dc", dc, "max rel diff before fail %.2e" % d.max(), "first >1e-6 at", bad[:3], "t there", tt[i, bad[:1]], "t_fail", tt[i, k - 1], "oracle nrhs", sto.nrhs)
