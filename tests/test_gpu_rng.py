"""The device's counter-based RNG (csrc/philox.cuh) pinned element by element: raw Philox4x32-10 against the
Random123 known answers and the oracle's C restatement; the normal / uniform / gamma draws of a (seed, chain,
iteration) against the oracle's restatement of the documented stream layout (which tests/test_oracle_golden.py
pins against the laws themselves); and the draws the MCMC kernels report against the same hook."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPE = 0.5 * (0.01 + 500)          # (n0 + N)/2, MCMC.py:158


def _draws(torch, pkg, seed, id0, c, iter0, n_iters, shape=SHAPE):
    lib = pkg._lib.load()
    out = torch.empty((n_iters, 6, c), dtype=torch.float64, device="cuda")
    pkg._lib.check(lib.rsfm_philox_draws(seed, id0, c, iter0, n_iters, shape, out.data_ptr(), None), "rsfm_philox_draws")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_device_philox_known_answers(cuda, pkg, orc):
    torch = cuda
    lib = pkg._lib.load()
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]          # Random123 kat_vectors, philox4x32-10
    rng = np.random.default_rng(0)
    rand = rng.integers(0, 2 ** 32, size=(500, 6), dtype=np.uint64).astype(np.uint32)
    inp = np.concatenate([np.array([c + k for c, k, _ in kat], dtype=np.uint32), rand])
    t_in = torch.from_numpy(inp.view(np.int32).copy()).cuda()
    t_out = torch.empty((inp.shape[0], 4), dtype=torch.int32, device="cuda")
    pkg._lib.check(lib.rsfm_philox_raw(t_in.data_ptr(), t_out.data_ptr(), inp.shape[0], None), "rsfm_philox_raw")
    out = t_out.cpu().numpy().view(np.uint32)
    for i, (_, _, want) in enumerate(kat):
        assert list(out[i]) == want
    for i in range(3, inp.shape[0]):
        assert np.array_equal(out[i], orc.philox4x32_10(inp[i, :4], inp[i, 4:]))


@pytest.mark.parametrize("seed,id0", [(20240, 0), (0xDEADBEEFCAFE, 1000), (7, (1 << 32) - 3), (2 ** 63 + 11, 2 ** 40)])
def test_device_draws_match_the_oracle_element_by_element(cuda, pkg, orc, seed, id0):
    """(z0, z1, z2, U, gamma, attempts) for 96 chains x 7 iterations: the uniforms are the same 53 bits, the
    normals agree to 1e-15 (libm vs CUDA log / sincospi), the gammas to 1e-14 relative with the same number of
    Marsaglia-Tsang attempts.  Chain ids beyond 2^32 use both counter words; seeds beyond 2^32 both key words."""
    c, iter0, n = 96, 123456, 7
    dev = _draws(cuda, pkg, seed, id0, c, iter0, n)
    ref = np.array([[orc.philox_draws(seed, id0 + ch, iter0 + i, SHAPE) for ch in range(c)] for i in range(n)])   # [n, c, 6]
    ref = ref.transpose(0, 2, 1)
    assert np.array_equal(dev[:, 3], ref[:, 3])                                  # U: bit-exact
    assert np.max(np.abs(dev[:, :3] - ref[:, :3])) <= 1e-15 * max(1.0, np.max(np.abs(ref[:, :3])))
    assert np.array_equal(dev[:, 5], ref[:, 5])
    assert np.max(np.abs(dev[:, 4] / ref[:, 4] - 1)) <= 1e-14


def test_device_gamma_and_normal_laws(cuda, pkg):
    """Kolmogorov-Smirnov and moments of the device draws at the sampler's shape 250.005 (and a small shape that
    exercises the rejection loop)."""
    from scipy import stats
    d = _draws(cuda, pkg, 31337, 5000, 4096, 0, 50)                 # 204,800 draws of each kind
    z = d[:, :3].reshape(-1)
    assert stats.kstest(z[::3], "norm").pvalue > 1e-3
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 4 * np.sqrt(2 / z.size)
    assert abs(np.corrcoef(d[:, 0].reshape(-1), d[:, 1].reshape(-1))[0, 1]) < 4 / np.sqrt(d[:, 0].size)
    u = d[:, 3].reshape(-1)
    assert stats.kstest(u, "uniform").pvalue > 1e-3 and u.min() > 0 and u.max() < 1
    g = d[:, 4].reshape(-1)
    assert stats.kstest(g, "gamma", args=(SHAPE,)).pvalue > 1e-3
    assert abs(g.mean() - SHAPE) < 4 * np.sqrt(SHAPE / g.size)
    assert abs(g.var() / SHAPE - 1) < 4.4 * np.sqrt(2 / g.size)
    assert abs(stats.skew(g) - 2 / np.sqrt(SHAPE)) < 5 * np.sqrt(6 / g.size)
    assert 1.0 <= d[:, 5].mean() < 1.05
    g2 = _draws(cuda, pkg, 1, 0, 4096, 0, 10, shape=1.5)[:, 4].reshape(-1)
    assert stats.kstest(g2, "gamma", args=(1.5,)).pvalue > 1e-3


@pytest.mark.parametrize("spec_depth", [1, 0])
def test_kernel_draws_are_the_documented_stream(cuda, pkg, spec_depth):
    """What rsfm_run reports as the draws it used (proposal, U, unit gamma) is exactly the (seed, global chain
    id, iteration) stream of the hook: proposal = q + sqrt(V) z0, U and gamma bit for bit -- for the sequential
    and the speculative kernel, across two launches (the iteration counter carries on)."""
    from conftest import load_golden
    torch = cuda
    g = load_golden("sse_grid.json")
    lib = pkg._lib.load()
    cfg = pkg.RateStateModel().to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = 1, 3, spec_depth
    cfg.lo[0], cfg.hi[0] = 1200.0, 1500.0
    c, seed, id0 = 40, 99, 77777
    q0 = torch.full((1, c), 1320.0, dtype=torch.float64, device="cuda")
    data_t = torch.from_numpy(g["data"]).cuda()
    h = lib.rsfm_create(C.byref(cfg), c, seed, id0)
    assert h
    try:
        pkg._lib.check(lib.rsfm_init(h, q0.data_ptr(), data_t.data_ptr(), None))
        var = torch.empty((1, c), dtype=torch.float64, device="cuda")
        pkg._lib.check(lib.rsfm_get_state(h, None, None, None, var.data_ptr(), None, None, None, None, None))
        rows = []
        for ns in (9, 6):
            samples = torch.empty((ns, 1, c), dtype=torch.float64, device="cuda")
            draws = torch.empty((ns, 3, c), dtype=torch.float64, device="cuda")
            pkg._lib.check(lib.rsfm_run(h, ns, samples.data_ptr(), None, None, draws.data_ptr(), None))
            torch.cuda.synchronize()
            rows.append((samples.cpu().numpy(), draws.cpu().numpy()))
    finally:
        lib.rsfm_destroy(h)
    samples = np.concatenate([r[0] for r in rows])[:, 0]            # [15, c]
    draws = np.concatenate([r[1] for r in rows])                    # [15, 3, c]
    hook = _draws(torch, pkg, seed, id0, c, 0, 15)
    sd = np.sqrt(var.cpu().numpy()[0])
    cur = np.vstack([np.full((1, c), 1320.0), samples[:-1]])
    assert np.allclose(draws[:, 0], cur + sd * hook[:, 0], rtol=1e-15, atol=0)
    inb = ~np.isnan(draws[:, 1])
    assert np.array_equal(inb, (draws[:, 0] > 1200.0) & (draws[:, 0] < 1500.0)) and not inb.all() and inb.any()
    assert np.array_equal(draws[:, 1][inb], hook[:, 3][inb])
    assert np.array_equal(draws[:, 2], hook[:, 4])
