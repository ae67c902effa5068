"""Predictive speculation (rsf_mcmc_spec_kernel): time per iteration and executed / decided solves for a few shapes.
usage: python profiles/microbench/spec_predict.py"""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
lib = pkg._lib.load()
m = pkg.RateStateModel(); m.Dc = 1325.0
np.random.seed(2024)
_, _, data = m.evaluate()
data_t = torch.from_numpy(np.ascontiguousarray(data)).cuda()


def run(c, depth, iters, launches, far=False, compat=False):
    cfg = m.to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = 1, (2 if compat else 3), depth
    cfg.adapt_mode = pkg._lib.ADAPT_COMPAT if compat else pkg._lib.ADAPT_NONE
    cfg.lo[0], cfg.hi[0] = 0.0, 1e4
    rng = np.random.default_rng(1)
    q0 = rng.uniform(200.0, 5000.0, c) if far else np.full(c, 1000.0)
    q0_t = torch.from_numpy(q0.reshape(1, c).copy()).cuda()
    h = lib.rsfm_create(C.byref(cfg), c, 11, 0)
    assert h, lib.rsfm_last_error()
    try:
        pkg._lib.check(lib.rsfm_init(h, q0_t.data_ptr(), data_t.data_ptr(), None))
        g = lib.rsfm_spec_depth(h)
        samples = torch.empty((iters, 1, c), dtype=torch.float64, device="cuda")
        ts = []
        for _ in range(launches):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pkg._lib.check(lib.rsfm_run(h, iters, samples.data_ptr(), None, None, None, None))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        tot = (C.c_uint64 * 9)()
        pkg._lib.check(lib.rsfm_get_totals(h, tot, None))
    finally:
        lib.rsfm_destroy(h)
    decided, execd = tot[0] - 3 * c, tot[6] - 3 * c
    print(f"C={c:6d} depth={depth} (g={g}) iters/launch={iters:5d} far={int(far)} compat={int(compat)}: "
          f"{np.mean(ts[1:]) if launches > 1 else ts[0]:8.2f} ms/launch ({np.mean(ts[1:] if launches > 1 else ts) / iters * 1e3:7.1f} us/iter), "
          f"first {ts[0]:8.2f} ms; executed/decided {execd / max(1, decided):.3f}, "
          f"{c * iters / (np.mean(ts[1:] if launches > 1 else ts) * 1e-3) / 1e6:7.2f} M chain-iterations/s, mean {samples[iters // 2:].mean().item():.1f}")


for c, depth in ((1024, 4), (1024, 5), (1024, 3), (2048, 4), (2048, 3), (4096, 3), (4096, 2), (512, 5), (1, 5), (1, 4), (16, 5)):
    run(c, depth, 200, 6)
run(1024, 4, 2000, 2)
run(1024, 5, 2000, 2)
run(1024, 4, 200, 6, far=True)
run(1024, 5, 200, 6, far=True)
run(1, 5, 500, 3)
run(1, 5, 500, 3, compat=True)
