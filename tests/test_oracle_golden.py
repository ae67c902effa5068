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
