"""The CPU oracle (oracle/rsf_oracle.c) against golden vectors generated from the
UNMODIFIED reference by oracle/make_golden.py.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden

# Oracle vs reference on the same machine differ only through libm-vs-NumPy ulps in
# exp/log/sin (and are bit-identical for most Dc).  Gate: 1e-12 relative to max|acc|.
TRAJ_RTOL = 1e-12
# stiff regime: accepted-step sequences legitimately diverge
TRAJ_RTOL_STIFF = 1e-5


def test_philox_known_answers(orc):
    # Random123 kat_vectors for philox4x32-10
    assert list(orc.philox4x32_10([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(orc.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6,
                                                                          0x6d5451fd]
    assert list(orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                  [0xa4093822, 0x299f31d0])) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_forward_trajectories_match_reference(orc):
    g = load_golden("forward_trajectories.json")
    assert g["_stamp"]["scipy"]          # oracle identity = SciPy version (the reference pins none)
    n_exact = 0
    for case in g["cases"]:
        kw = {}
        if "end_time" in case:
            kw["end_time"] = case["end_time"]
        m = orc.make_model(Dc=case["Dc"], number_time_steps=case["N"],
                           radiation_damping=int(case["RadiationDamping"]), **kw)
        t, acc, st = orc.forward(m)
        ref = case["acc"]
        assert acc.shape == ref.shape
        if "filled" in case:
            # silent-failure case (quirk q9): same number of filled entries, zero tail
            assert st.filled == case["filled"] and st.istate == -2
            assert np.all(acc[st.filled:] == 0.0)
            assert t[1] == pytest.approx(case["t1"], rel=1e-4)
            continue
        scale = np.max(np.abs(ref))
        tol = TRAJ_RTOL_STIFF if case["Dc"] < 1.0 else TRAJ_RTOL
        assert np.max(np.abs(acc - ref)) <= tol * scale, case["Dc"]
        assert t[-1] == case["t_last"]
        n_exact += int(np.array_equal(acc, ref))
    assert n_exact >= 6     # most non-stiff cases are reproduced bit for bit


def test_spot_values_from_survey(orc):
    # SURVEY.md 8c G1 spot values, Dc = 1000, damping on, N = 500
    _, acc, st = orc.forward(orc.make_model(Dc=1000.0))
    assert acc[1] == 0.00416511121489771
    assert acc[2] == 0.008630414475125914
    assert acc[10] == -0.0006357780297205906
    assert acc[100] == -0.004554832082448446
    assert acc[250] == -0.0024297302243581065
    assert acc[499] == 0.0006071259074835744
    assert st.nrhs == 7034 and st.nstep == 503


def test_sse_grid(orc):
    g = load_golden("sse_grid.json")
    m = orc.make_model()
    sse, _, _ = orc.forward_batch(m, g["grid"], data=g["data"])
    assert np.allclose(sse, g["sse"], rtol=1e-12, atol=0)
    assert g["grid"][int(np.argmin(sse))] == 1300.0        # grid minimum at 1300


@pytest.mark.parametrize("name", ["chain_list_priors.json", "chain_dict_priors.json", "chain_bounds.json", "chain_cfg1_full.json"])
def test_chain_replay_bit_exact(orc, name):
    """Appendix A semantics: with the recorded draws the oracle reproduces the reference chain."""
    g = load_golden(name)
    ns, nb = g["nsamples"], g["nburn"]
    m = orc.make_model()
    uni = np.nan_to_num(g["uniforms"], nan=0.5)       # NaN = U not drawn (out-of-bounds proposal, q10)
    chain, s2, acc, vstart, nsolves = orc.chain_replay(m, g["data"], g["qstart"], g["lo"], g["hi"],
                                                       g["n_prior_len"], ns, g["proposals"], uni, g["gammas_unit"])
    assert np.array_equal(chain[nb:], g["chain_post_burn"][0])
    assert np.array_equal(acc, g["accepts"])
    assert np.allclose(s2[nb:], g["std2_post_burn"], rtol=1e-13, atol=0)
    assert vstart == pytest.approx(g["Vstart"].item(), rel=1e-13)
    assert nsolves == 3 + int(np.sum(~np.isnan(g["uniforms"])))


def test_compat_adaptation_matches_reference(orc):
    """dict-typed priors: the oracle's own adaptation (quirk q3) reproduces the proposal scales used."""
    g = load_golden("chain_dict_priors.json")
    ns = g["nsamples"]
    m = orc.make_model()
    uni = np.nan_to_num(g["uniforms"], nan=0.5)
    full, _, _, _, _ = orc.chain_replay(m, g["data"], g["qstart"], g["lo"], g["hi"], g["n_prior_len"], ns,
                                        g["proposals"], uni, g["gammas_unit"])
    z = (g["proposals"] - full[:-1]) / np.sqrt(g["V_used"])
    chain2, _, acc2, _, _ = orc.chain_replay(m, g["data"], g["qstart"], g["lo"], g["hi"], g["n_prior_len"], ns, z,
                                             uni, g["gammas_unit"], compat_adapt=True)
    assert np.array_equal(acc2, g["accepts"])
    assert np.allclose(chain2, full, rtol=1e-12, atol=0)
    assert len(np.unique(g["V_used"])) > 3            # the scale really changed during the run


def test_list_priors_never_adapt():
    g = load_golden("chain_list_priors.json")
    assert len(np.unique(g["V_used"])) == 1           # quirk q2


# ---- the NumPy/SciPy form of the oracle (oracle/scipy_port.py), the CPU baseline bench.py times ----
def test_scipy_port_trajectory_bit_exact():
    from oracle import scipy_port
    g = load_golden("forward_trajectories.json")
    case = [c for c in g["cases"] if c["Dc"] == 1350.0 and c["RadiationDamping"]][0]
    m = scipy_port.PortModel()
    m.Dc = 1350.0
    t, acc, _ = m.evaluate()
    assert np.array_equal(acc, case["acc"])
    assert m.n_rhs >= 7034


def test_scipy_port_chain_bit_exact():
    """Same seed, same draw order (normal -> [randn(N) -> rand] -> gamma) => the reference's chain."""
    from oracle import scipy_port
    g = load_golden("chain_list_priors.json")
    ns, nb = 40, g["nburn"]
    r = scipy_port.run_chain(g["data"], g["qstart"], g["lo"], g["hi"], ns, n_prior_len=3, seed=g["seed"])
    # the golden chain has 120 iterations; the first 40 decisions and states must coincide
    assert np.array_equal(r["accepts"], g["accepts"][:ns].astype(bool))
    full_ref, _, _, _, _ = __import__("oracle.oracle", fromlist=["x"]).chain_replay(
        __import__("oracle.oracle", fromlist=["x"]).make_model(), g["data"], g["qstart"], g["lo"], g["hi"], 3,
        g["nsamples"], g["proposals"], np.nan_to_num(g["uniforms"], nan=0.5), g["gammas_unit"])[0:5]
    assert np.array_equal(r["chain"], full_ref[:ns + 1])
    assert r["n_solves"] == 3 + ns


def test_philox_samplers_of_the_oracle_have_the_right_laws(orc):
    """The oracle's restatement of the CUDA path's draw layout (normals by Box-Muller, 53-bit uniforms,
    Marsaglia-Tsang gamma at the sampler's shape (n0 + N)/2 = 250.005) is pinned here against the laws
    themselves -- Kolmogorov-Smirnov against N(0,1), U(0,1), Gamma(250.005) and their first two moments --
    so that the element-by-element comparison of the DEVICE draws with this oracle (tests/test_gpu_rng.py)
    also pins the distributions: a wrong counter layout or a biased gamma cannot pass both."""
    from scipy import stats
    shape = 0.5 * (0.01 + 500)
    n = 40000
    d = np.array([orc.philox_draws(987654321, 1000 + (i % 200), i // 200, shape) for i in range(n)])
    z = np.concatenate([d[:, 0], d[:, 1], d[:, 2]])
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 4 * np.sqrt(2 / z.size)
    assert abs(np.corrcoef(d[:, 0], d[:, 1])[0, 1]) < 4 / np.sqrt(n)
    assert stats.kstest(d[:, 3], "uniform").pvalue > 1e-3
    assert np.all((d[:, 3] > 0) & (d[:, 3] < 1))
    g = d[:, 4]
    assert stats.kstest(g, "gamma", args=(shape,)).pvalue > 1e-3
    assert abs(g.mean() - shape) < 4 * np.sqrt(shape / n)
    assert abs(g.var() / shape - 1) < 4 * np.sqrt(2 / n) * 1.1
    assert np.all(d[:, 5] >= 1) and 1.0 <= d[:, 5].mean() < 1.05        # squeeze accepts > 95 % of the attempts
    # small shapes too (the rejection loop is exercised): Gamma(1.5)
    g2 = np.array([orc.philox_draws(5, i, 0, 1.5)[4] for i in range(20000)])
    assert stats.kstest(g2, "gamma", args=(1.5,)).pvalue > 1e-3
    # streams: distinct chains / iterations give distinct draws, same counter gives the same draw
    assert np.array_equal(orc.philox_draws(7, 3, 9, shape), orc.philox_draws(7, 3, 9, shape))
    assert orc.philox_draws(7, 3, 9, shape)[0] != orc.philox_draws(7, 4, 9, shape)[0]
    assert orc.philox_draws(7, 3, 9, shape)[0] != orc.philox_draws(7, 3, 10, shape)[0]
    assert orc.philox_draws(7, 3, 9, shape)[0] != orc.philox_draws(8, 3, 9, shape)[0]
    assert orc.philox_draws(7, 3 + (1 << 32), 9, shape)[0] != orc.philox_draws(7, 3, 9, shape)[0]


def test_nd_replay_reduces_to_the_reference_chain_for_d1(orc):
    """orc_chain_replay_nd (absolute proposals, per-parameter bounds, d = 1 or 3) is the same loop as the
    d = 1 replay that reproduces the recorded reference chains bit for bit."""
    g = load_golden("chain_bounds.json")
    ns = g["nsamples"]
    uni = np.nan_to_num(g["uniforms"], nan=0.5)
    chain, s2, acc, _, nsolves = orc.chain_replay(orc.make_model(), g["data"], g["qstart"], g["lo"], g["hi"],
                                                  g["n_prior_len"], ns, g["proposals"], uni, g["gammas_unit"])
    chain_n, s2_n, acc_n, nsolves_n = orc.chain_replay_nd(orc.make_model(), g["data"], [g["qstart"]], [g["lo"]], [g["hi"]],
                                                         g["n_prior_len"], ns, g["proposals"], uni, g["gammas_unit"])
    assert np.array_equal(acc_n, acc) and np.array_equal(acc_n, g["accepts"])
    assert np.array_equal(chain_n[:, 0], chain) and np.array_equal(s2_n, s2)
    assert nsolves_n == nsolves - 2                    # no covariance set-up solves in the nd form
    # d = 3: proposals outside any one bound are rejected without a solve; the (a, b) of an accepted
    # proposal are used by the next solve
    rng = np.random.default_rng(0)
    q0 = np.array([0.011, 0.014, 1300.0])
    lo, hi = np.array([0.0105, 0.0135, 1200.0]), np.array([0.0115, 0.0145, 1400.0])
    props = q0 + rng.standard_normal((12, 3)) * np.array([2e-4, 2e-4, 40.0])
    props[3, 0] = 0.02
    ch3, s23, acc3, ns3 = orc.chain_replay_nd(orc.make_model(), g["data"], q0, lo, hi, 3, 12, props,
                                              rng.random(12), rng.gamma(250.005, size=12))
    inb = np.all((props > lo) & (props < hi), axis=1)
    assert ns3 == 1 + inb.sum() and not acc3[3] and acc3.sum() > 0
    for i in range(12):
        assert np.array_equal(ch3[i + 1], props[i] if acc3[i] else ch3[i])


def test_rhs_matches_reference_friction(orc):
    """SURVEY section 4, unit level: the RHS restatement against values of the reference's own nested
    ``friction(t, y)`` (captured unmodified by oracle/make_golden.py), both damping modes, near and far from
    sliding steady state: agreement to 1e-15 of the magnitude of the terms each component is a difference of."""
    from conftest import rhs_term_scales
    g = load_golden("rhs_values.json")
    assert g["columns"][:2] == ["RadiationDamping", "Dc"] and g["rows"].shape[1] == 9
    worst = 0.0
    for row in g["rows"]:
        m = orc.make_model(Dc=row[1], radiation_damping=int(row[0]))
        out = orc.rhs(m, row[2], row[3:6])
        worst = max(worst, float(np.max(np.abs(out - row[6:9]) / rhs_term_scales(row))))
    assert worst <= 1e-15, worst


# ---- k1 as the varied constant (SURVEY 8f.4 "named parameters"): golden from the unmodified reference with its
# ---- public attribute model.k1 changed (oracle/make_golden.py k1) ----
def test_forward_with_k1_matches_reference(orc):
    g = load_golden("forward_k1.json")
    n_exact = 0
    for case in g["cases"]:
        m = orc.make_model(Dc=case["Dc"], k1=case["k1"], number_time_steps=case["N"])
        t, acc, st = orc.forward(m)
        ref = case["acc"]
        assert np.max(np.abs(acc - ref)) <= TRAJ_RTOL * np.max(np.abs(ref)), (case["Dc"], case["k1"])
        assert t[-1] == case["t_last"]
        n_exact += int(np.array_equal(acc, ref))
    assert n_exact >= 12
    # k1 matters: the largest value roughly quarters the response
    big = [c for c in g["cases"] if c["Dc"] == 1000.0]
    assert np.max(np.abs(big[-1]["acc"])) < 0.3 * np.max(np.abs(big[0]["acc"]))


def test_sse_on_a_k1_grid_matches_reference(orc):
    """ORC_PARAM_K1: the batch axis (and the chain's scalar) is k1, Dc stays fixed."""
    g = load_golden("forward_k1.json")
    m = orc.make_model(Dc=g["data_Dc"], sampled_param=orc.PARAM_K1)
    sse, _, _ = orc.forward_batch(m, g["k1_grid"], data=g["data"])
    assert np.allclose(sse, g["sse"], rtol=1e-11, atol=0)
    assert g["k1_grid"][int(np.argmin(g["sse"]))] in (2.5e-3, 3e-3, 3.5e-3)     # truth 3e-3


def test_scipy_port_with_k1_bit_exact():
    from oracle import scipy_port
    g = load_golden("forward_k1.json")
    case = [c for c in g["cases"] if c["Dc"] == 1350.0 and c["k1"] == 3e-3][0]
    m = scipy_port.PortModel()
    m.Dc, m.k1 = 1350.0, 3e-3
    assert np.array_equal(m.evaluate()[1], case["acc"])


def test_k1_replay_is_the_dc_replay_with_the_attribute_swapped(orc):
    """orc_chain_replay with ORC_PARAM_K1 against a plain Python loop over orc.forward (MCMC.py:245-266, 494-521 with
    model.k1 in the place of model.Dc)."""
    g = load_golden("forward_k1.json")
    data = g["data"]
    rng = np.random.default_rng(7)
    ns, q0, lo, hi = 12, 2e-3, 0.0, 0.01
    z, u, gam = rng.standard_normal(ns), rng.random(ns), rng.gamma(0.5 * (0.01 + 500), size=ns)
    m = orc.make_model(Dc=1000.0, sampled_param=orc.PARAM_K1)
    chain, s2, acc, vstart, nsolves = orc.chain_replay(m, data, q0, lo, hi, 3, ns, z, u, gam, compat_adapt=True)

    def sse_of(k1):
        return orc.sse(orc.forward(orc.make_model(Dc=1000.0, k1=k1))[1], data)
    base = orc.forward(orc.make_model(Dc=1000.0, k1=q0))[1]
    pert = orc.forward(orc.make_model(Dc=1000.0, k1=q0 * (1 + 1e-6)))[1]
    s2_0 = orc.sse(base, data) / (500 - 3)
    x = (pert - base) / (q0 * (1 + 1e-6) * 1e-6)
    v = s2_0 / float(np.sum(x * x))
    assert vstart == pytest.approx(v, rel=1e-12)
    q, ss, s2_i = q0, sse_of(q0), s2_0
    for i in range(ns):
        qn = q + np.sqrt(v) * z[i]
        ok = lo < qn < hi
        if ok:
            ssn = sse_of(qn)
            ok = min(0.0, 0.5 * (ss - ssn) / s2_i) > np.log(u[i])
            if ok:
                q, ss = qn, ssn
        assert bool(acc[i]) == bool(ok)
        assert chain[i + 1] == q
        s2_i = 1 / (gam[i] * (1 / (0.5 * (0.01 * s2_i + ss))))
        assert s2[i + 1] == pytest.approx(s2_i, rel=1e-13)
        if (i + 1) % 10 == 0:                                        # MCMC.py:523-527, 200-204 (dict-prior form)
            vnew = 2.38 ** 2 / 2.0 * np.var(chain[i + 2 - 10:i + 2], ddof=1)
            if vnew > 0:
                v = np.sqrt(vnew)
