"""Predictive speculation, joint (a, b, Dc) posterior (d = 3: ten-coefficient fit): time per iteration and executed /
decided solves with the speculative kernel against one thread per chain.
usage: python profiles/microbench/spec_predict3.py"""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
lib = pkg._lib.load()
m = pkg.RateStateModel(); m.Dc = 1325.0
np.random.seed(2024)
_, _, data = m.evaluate()
data_t = torch.from_numpy(np.ascontiguousarray(data)).cuda()


def run(c, depth, iters, launches, pooled=False):
    cfg = m.to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = 3, 3, depth
    cfg.adapt_mode = pkg._lib.ADAPT_NONE
    for j, (lo, hi) in enumerate(((0.005, 0.02), (0.005, 0.03), (0.0, 1e4))):
        cfg.lo[j], cfg.hi[j] = lo, hi
    rng = np.random.default_rng(1)
    q0 = np.stack([rng.uniform(0.0105, 0.0115, c), rng.uniform(0.0135, 0.0145, c), rng.uniform(1000.0, 1800.0, c)], axis=0)
    q0_t = torch.from_numpy(np.ascontiguousarray(q0)).cuda()
    h = lib.rsfm_create(C.byref(cfg), c, 11, 0)
    assert h, lib.rsfm_last_error()
    try:
        pkg._lib.check(lib.rsfm_init(h, q0_t.data_ptr(), data_t.data_ptr(), None))
        g = lib.rsfm_spec_depth(h)
        samples = torch.empty((iters, 3, c), dtype=torch.float64, device="cuda")
        acc = torch.empty((iters, c), dtype=torch.uint8, device="cuda")
        ts = []
        for _ in range(launches):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pkg._lib.check(lib.rsfm_run(h, iters, samples.data_ptr(), None, acc.data_ptr(), None, None))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        tot = (C.c_uint64 * 9)()
        pkg._lib.check(lib.rsfm_get_totals(h, tot, None))
    finally:
        lib.rsfm_destroy(h)
    decided, execd = tot[0] - 4 * c, tot[6] - 4 * c
    t = np.mean(ts[1:]) if launches > 1 else ts[0]
    print(f"d=3 C={c:5d} depth={depth} (g={g}) iters/launch={iters:4d}: {t:8.2f} ms/launch ({t / iters * 1e3:7.1f} us/iter), first {ts[0]:8.2f} ms; "
          f"executed/decided {execd / max(1, decided):.3f}, accept {acc.float().mean().item():.2f}, {c * iters / (t * 1e-3) / 1e6:6.2f} M chain-iterations/s")


for c, depth in ((1, 0), (1, 1), (64, 0), (1024, 0), (1024, 1), (2048, 0), (4096, 0), (4096, 1)):
    run(c, depth, 200, 4)
