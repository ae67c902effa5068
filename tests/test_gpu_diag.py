"""Device diagnostics and post-processing against NumPy / SciPy definitions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ess_numpy(x, max_lag):
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    e = x - x.mean()
    c0 = float(e @ e)
    tau, prev, t = 0.0, 1e300, 0
    while t + 1 <= min(max_lag, n - 1):
        pair = (float(e[:n - t] @ e[t:]) + float(e[:n - t - 1] @ e[t + 1:])) / c0
        if pair <= 0:
            break
        pair = min(pair, prev)
        prev = pair
        tau += 2 * pair
        t += 2
    return n / max(tau - 1.0, 1.0 / n)


def test_chain_diagnostics_match_numpy(cuda, pkg):
    import importlib
    torch = cuda
    dg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200.diagnostics")
    rng = np.random.default_rng(0)
    n, d, c = 400, 2, 37
    x = np.zeros((n, d, c))
    for j in range(c):                                     # AR(1) chains with different correlation
        rho = 0.1 + 0.8 * j / c
        for p in range(d):
            e = rng.standard_normal(n)
            for i in range(1, n):
                x[i, p, j] = rho * x[i - 1, p, j] + e[i]
    out = dg.chain_diagnostics(torch.from_numpy(x).cuda(), max_lag=200)
    for p in range(d):
        ess = sum(_ess_numpy(x[:, p, j], 200) for j in range(c))
        assert out["ess"][p] == pytest.approx(ess, rel=1e-9)
        half = n // 2
        halves = np.concatenate([x[:half, p, :], x[n - half:, p, :]], axis=1)        # [half, 2c]
        w = halves.var(axis=0, ddof=1).mean()
        b_over_n = halves.mean(axis=0).var(ddof=1)
        assert out["rhat"][p] == pytest.approx(np.sqrt(((half - 1) / half * w + b_over_n) / w), rel=1e-9)
        assert out["mean"][p] == pytest.approx(x[:, p, :].mean(), abs=1e-12)
        assert out["sd"][p] == pytest.approx(x[:, p, :].std(), rel=1e-9)


def test_kde_matches_scipy(cuda, pkg):
    from scipy.stats import gaussian_kde
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.normal(1340.0, 60.0, 700), rng.normal(1500.0, 20.0, 300)])
    grid, pdf = pkg.gaussian_kde_pdf(x, lo=1100.0, hi=1700.0)
    assert grid.shape == (1000,) and pdf.shape == (1000,)
    ref = gaussian_kde(x).pdf(grid)
    assert np.allclose(pdf, ref, rtol=1e-10, atol=1e-300)
    assert np.trapezoid(pdf, grid) == pytest.approx(1.0, abs=2e-3)
    with pytest.raises(np.linalg.LinAlgError):
        pkg.gaussian_kde_pdf(np.ones(10))
