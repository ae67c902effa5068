"""Which path do the steps of the stiff regime take?  Needs a counting build of the library:
    nvcc ... -DRSFM_DEBUG_COUNT -o librsfm_dbg.so ... ; RSFM_LIB=$PWD/librsfm_dbg.so python profiles/microbench/stiff_paths.py
(tuning tool only; the product library has no counters)."""
import ctypes, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
lib = ctypes.CDLL(os.environ["RSFM_LIB"])
n = 4000
m = pkg.RateStateModel(number_time_steps=n, end_time=n * 0.1)
m.loading, m.vstep_period, m.vstep_factor = "vstep", 100.0, 10.0
names = ["lane-steps", "lane-steps fast tried", "lane-steps general", "warp-steps", "warp-steps general", "warp-steps fast tried",
         "lane-steps accepted", "warp fast intervals", "lane-steps raised level", "lane-steps out of range from an in-range start", "lane rejected + out of range"]
for label, dc_h in (("all 0.05", np.full(2048, 0.05)), ("U(0.03,0.08)", np.random.default_rng(0).uniform(0.03, 0.08, 2048))):
    buf = (ctypes.c_ulonglong * 16)()
    lib.rsfm_debug_counters(buf, 1)
    out = m.evaluate_batch(torch.from_numpy(dc_h).cuda(), want_acc=False, data=np.zeros(m.num_outputs()))
    torch.cuda.synchronize()
    lib.rsfm_debug_counters(buf, 1)
    print(label, "nstep", int(out["nstep"].sum()))
    for i, nm in enumerate(names):
        print(f"   {nm:28s} {buf[i]}")
