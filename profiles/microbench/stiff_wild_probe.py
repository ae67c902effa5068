"""Which steps does the stiff variant's 'presumed wild' rule (rsfm_device.cuh, rsf_interval_general) cover that the
exact arithmetic accepts?  Debug build only (-DRSFM_DEBUG_COUNT -> librsfm_dbg.so): in stiff_exact mode every step
that starts inside half the fast ranges, leaves them at a stage and is ACCEPTED by the general-range stages is recorded.

Build of the counting library (next to the product's, never loaded by it):
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -DRSFM_DEBUG_COUNT \
         -I include -I bayesian-markov-chain-monte-carlo_b200/csrc \
         -o bayesian-markov-chain-monte-carlo_b200/librsfm_dbg.so bayesian-markov-chain-monte-carlo_b200/csrc/rsfm_kernels.cu
Round-2 result (before the rule got its load condition): one such step per velocity jump -- the first step behind a
jump that the accumulated output times put a few ulp inside an interval whose frame is the old level (t = 90.000132,
h = 6.5e-4, err = 4.6e-13 at Dc = 2) -- and 1.3e-6 of the trajectory between the rule and the exact arithmetic; with
the load condition the two give the same bits (tests/test_gpu_forward.py::test_stiff_rule_changes_no_decision)."""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
pkg._lib.LIB_PATH = os.path.join(os.path.dirname(pkg._lib.LIB_PATH), "librsfm_dbg.so")
lib = pkg._lib.load()
import torch

def records():
    out = (C.c_double * 512)(); n = C.c_uint(0)
    lib.rsfm_debug_records(out, C.byref(n), 1)
    cnt = (C.c_ulonglong * 16)()
    lib.rsfm_debug_counters(cnt, 1)
    return np.array(out).reshape(64, 8)[:min(n.value, 64)], n.value, list(cnt)

for (n, t_end, period, factor, dcs) in [(1200, 120.0, 30.0, 10.0, [2.0]), (1200, 120.0, 30.0, 10.0, [0.05, 0.3, 0.6, 1.0, 1.5, 3.0, 5.0, 10.0]),
                                         (600, 60.0, 20.0, 3.0, [0.05, 0.3, 1.0, 2.0, 5.0])]:
    for dc in dcs:
        m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
        m.loading, m.vstep_period, m.vstep_factor = "vstep", period, factor
        m.solver_variant = "stiff"
        res = {}
        for exact in (True, False):
            m.stiff_exact = exact
            records()
            o = m.evaluate_batch(np.array([dc]))
            rec, nrec, cnt = records()
            res[exact] = o["acc"].t().cpu().numpy()[0]
            if exact:
                print(f"Dc={dc} period={period} x{factor}: steps={cnt[0]} bad={cnt[2]} bad&start_in={cnt[9]} of which ACCEPTED by exact arithmetic={cnt[11]}")
                for r in rec[:6]:
                    print("    t=%.6f h=%.3e f0/lim=%.3f A0/lim=%.3f errA=%.3e den3=%.3e err=%.3e w=%.4g lam=%g" %
                          (r[0], r[1], r[2], r[3], r[4], r[5], abs(r[1]) * r[4] / np.sqrt(r[5]), r[6], r[7]))
        print("    exact vs rule: max rel diff %.3e" % (np.max(np.abs(res[True] - res[False])) / np.max(np.abs(res[True]))))
