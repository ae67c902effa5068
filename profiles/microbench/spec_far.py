"""Which chains hold a far-start launch of the speculative kernel up?  Per-chain executed work after a few launches."""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
lib = pkg._lib.load()
m = pkg.RateStateModel(); m.Dc = 1325.0
np.random.seed(2024)
_, _, data = m.evaluate()
data_t = torch.from_numpy(np.ascontiguousarray(data)).cuda()
c, iters = 1024, 200
for depth in (4, 5):
    cfg = m.to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = 1, 3, depth
    cfg.lo[0], cfg.hi[0] = 0.0, 1e4
    rng = np.random.default_rng(1)
    q0 = rng.uniform(200.0, 5000.0, c)
    q0_t = torch.from_numpy(q0.reshape(1, c).copy()).cuda()
    h = lib.rsfm_create(C.byref(cfg), c, 11, 0)
    pkg._lib.check(lib.rsfm_init(h, q0_t.data_ptr(), data_t.data_ptr(), None))
    prev = None
    for l in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pkg._lib.check(lib.rsfm_run(h, iters, None, None, None, None, None))
        e1.record(); torch.cuda.synchronize()
        q = torch.empty((1, c), dtype=torch.float64, device="cuda"); ch = torch.empty((1, c), dtype=torch.float64, device="cuda")
        acc = torch.empty(c, dtype=torch.int32, device="cuda"); st = torch.empty(c, dtype=torch.int32, device="cuda")
        nrhs = torch.empty(c, dtype=torch.int64, device="cuda"); nstep = torch.empty(c, dtype=torch.int64, device="cuda")
        pkg._lib.check(lib.rsfm_get_state(h, q.data_ptr(), None, None, ch.data_ptr(), acc.data_ptr(), st.data_ptr(), nrhs.data_ptr(), nstep.data_ptr(), None))
        torch.cuda.synchronize()
        nr = nrhs.cpu().numpy().astype(np.float64); a = acc.cpu().numpy()
        d = nr - (prev[0] if prev else 0); da = a - (prev[1] if prev else 0)
        prev = (nr, a)
        sd = np.sqrt(ch.cpu().numpy()[0]); qq = q.cpu().numpy()[0]
        top = np.argsort(-d)[:4]
        print(f"depth {depth} launch {l}: {e0.elapsed_time(e1):7.2f} ms; RHS per chain: median {np.median(d):.3g}, max {d.max():.3g}; "
              f"accepts per chain median {np.median(da):.0f}; top: " + ", ".join(f"[q0 {q0[i]:.0f} q {qq[i]:.0f} sd {sd[i]:.0f} acc {da[i]} rhs {d[i]:.3g}]" for i in top))
        pairs = d.reshape(-1, 2).max(axis=1) if depth == 4 else d
        print(f"      per-warp max RHS: median {np.median(pairs):.3g}, p99 {np.percentile(pairs, 99):.3g}, max {pairs.max():.3g}; proposal sd: median {np.median(sd):.0f}, max {sd.max():.0f}")
    lib.rsfm_destroy(h)
