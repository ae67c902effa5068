"""The scheduling policy of the speculative kernel replayed on the CPU oracle: whatever tree a round
evaluates, the chain is the sequential one; with the quadratic fit of SS in 1/Dc nearly every evaluated node lies on
the realised path.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden


@pytest.fixture(scope="module")
def problem(orc):
    g = load_golden("sse_grid.json")
    rng = np.random.default_rng(3)
    n = 240
    return dict(model=orc.make_model(), data=g["data"], z=rng.standard_normal(n + 40), u=rng.random(n + 40),
                gam=rng.gamma(0.5 * (0.01 + 500), size=n + 40), n=n)


@pytest.mark.parametrize("q0, sd", [(1000.0, 60.0), (4000.0, 150.0)])
def test_any_tree_gives_the_sequential_chain_and_the_fit_predicts_the_decisions(orc, problem, q0, sd):
    from oracle import predictive_tree as pt
    p = problem
    n = p["n"]
    args = (p["model"], p["data"], q0, sd * sd, 0.0, 1e4, p["z"][:n], p["u"][:n], p["gam"][:n])
    ref_chain, ref_acc = pt.sequential_chain(*args)
    assert 0.1 < ref_acc.mean() < 0.9
    res = {}
    for lanes, use_fit in ((16, False), (16, True), (32, True), (2, True)):
        chain, acc, st = pt.speculative_chain(*args, lanes=lanes, use_fit=use_fit)
        assert np.array_equal(chain, ref_chain) and np.array_equal(acc, ref_acc), (lanes, use_fit)
        res[(lanes, use_fit)] = st
    # acceptance-rate trees advance like the balanced tree of round 1 (4 of 15 nodes at 16 lanes) ...
    assert 3.0 < res[(16, False)]["advance"] < 6.5
    # ... the fit puts nearly every node on the realised path
    assert res[(16, True)]["advance"] > 11.0 and res[(32, True)]["advance"] > 18.0 and res[(2, True)]["advance"] > 1.7
    assert res[(16, True)]["predicted"] > 0.95
    assert res[(16, True)]["evaluated"] < 0.5 * res[(16, False)]["evaluated"]


def test_sum_of_squares_is_quadratic_in_the_reciprocal_of_dc(orc, problem):
    """The premise of the predictor: over +-350 around the mode a quadratic in 1/Dc leaves a residual far below
    sigma^2 (= SS/(N - 3)), a quadratic in Dc does not."""
    p = problem
    qs = np.linspace(900.0, 1600.0, 57)
    ss = orc.forward_batch(p["model"], qs, data=p["data"], nthreads=4)[0]
    s2 = ss.min() / 497.0

    def resid(x):
        x = (x - x.mean()) / x.std()
        return np.max(np.abs(ss - np.polyval(np.polyfit(x, ss, 2), x))) / s2
    assert resid(1.0 / qs) < 0.05
    assert resid(qs) > 1.0


def test_joint_posterior_trees_with_the_ten_coefficient_fit(orc, problem):
    """d = 3: the policy with the quadratic in (1/a, b/a, 1/Dc): the sequential chain whatever the tree, most nodes
    on the realised path."""
    from oracle import predictive_tree as pt
    rng = np.random.default_rng(9)
    n = 200
    z, u, gam = rng.standard_normal((n, 3)), rng.random(n), rng.gamma(0.5 * (0.01 + 500), size=n)
    args = (problem["data"], [0.0105, 0.0145, 1200.0], [2e-4, 3e-4, 40.0], [0.005, 0.005, 0.0], [0.02, 0.03, 1e4], z, u, gam)
    ref_chain, ref_acc, _ = pt.chain_abdc(*args, lanes=0)
    assert 0.1 < ref_acc.mean() < 0.9
    res = {}
    for use_fit in (False, True):
        chain, acc, st = pt.chain_abdc(*args, lanes=16, use_fit=use_fit)
        assert np.array_equal(chain, ref_chain) and np.array_equal(acc, ref_acc)
        res[use_fit] = st
    assert res[False]["advance"] < 6.5
    assert res[True]["advance"] > 8.5 and res[True]["predicted"] > 0.9
    assert res[True]["evaluated"] < 0.6 * res[False]["evaluated"]
