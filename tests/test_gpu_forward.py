"""Parity of the CUDA forward solve (rsf_forward_kernel through the C ABI) with the CPU oracle
and with golden vectors from the unmodified reference."""
import numpy as np
import pytest

import json
import os

from conftest import ROOT, load_golden, oracle_ulp_floor

pytestmark = pytest.mark.gpu

# |acc_gpu - acc_ref| <= 1e-9 * max|acc_ref| + 1e-12 outside the stiff regime,
# 1e-6 * max|acc_ref| for Dc < 50 where accepted-step sequences may legitimately differ.
RTOL, ATOL, RTOL_STIFF = 1e-9, 1e-12, 1e-6
# Velocity-step loading in the stability-limited regime: the reference solver itself is reproducible only to its
# own tolerance there -- a ONE-ulp change of Dc moves the CPU oracle's trajectory by 2-4e-6 of its scale, and SciPy
# and the C oracle differ by the same amount (tests/test_oracle_vs_scipy.py) -- so the stated 1e-6 is below what any
# two correct implementations can agree to.  The gate is max(1e-6, STIFF_FLOOR_MULT x the floor measured on the
# oracle for the very case under test); the observed errors are recorded (gpurun_out/observed_parity.json).
STIFF_FLOOR_MULT = 4.0


def _record(name, **values):
    path = os.path.join(ROOT, "gpurun_out", "observed_parity.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        try:
            log = json.load(open(path))
        except (OSError, ValueError):
            log = {}
        log[name] = values
        json.dump(log, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def _check_stiff(orc, acc_g, dc, tag, **model_kw):
    """GPU trajectory against the oracle with the measured-floor gate; returns (err, floor), both relative."""
    acc_o, scale, floor = oracle_ulp_floor(orc, dc, **model_kw)
    err = float(np.max(np.abs(acc_g - acc_o))) / scale
    gate = max(RTOL_STIFF, STIFF_FLOOR_MULT * floor)
    _record(tag, Dc=dc, err_rel=err, oracle_ulp_floor_rel=floor, gate_rel=gate)
    assert err <= gate, f"{tag}: err {err:.3e} > gate {gate:.3e} (oracle floor {floor:.3e})"
    return err, floor


def _check(acc_g, acc_ref, dc):
    scale = np.max(np.abs(acc_ref))
    tol = (RTOL_STIFF if dc < 50 else RTOL) * scale + ATOL
    err = np.max(np.abs(acc_g - acc_ref))
    assert err <= tol, f"Dc={dc}: err {err:.3e} > tol {tol:.3e}"
    return err / scale


def test_golden_trajectories(cuda, pkg):
    g = load_golden("forward_trajectories.json")
    worst = 0.0
    for case in g["cases"]:
        if "filled" in case:
            continue
        m = pkg.RateStateModel(number_time_steps=case["N"], end_time=case.get("end_time", 50.0))
        m.RadiationDamping = case["RadiationDamping"]
        m.Dc = case["Dc"]
        t, acc, acc_noise = m.evaluate()
        assert acc.shape == case["acc"].shape and acc.dtype == np.float64
        assert t[-1] == pytest.approx(case["t_last"], abs=1e-12)
        assert acc[0] == 0.0
        rel = _check(acc, case["acc"], case["Dc"])
        if case["Dc"] >= 50:
            worst = max(worst, rel)
    assert worst < 1e-11      # expected agreement is ~1e-13; the gate above is 1e-9


def test_batch_vs_oracle_and_counters(cuda, pkg, orc):
    dcs = np.array([0.05, 0.5, 1.0, 10.0, 60.0, 100.0, 333.0, 1000.0, 1350.0, 2500.0, 5000.0, 9999.0, 20000.0])
    m = pkg.RateStateModel()
    out = m.evaluate_batch(dcs, want_t=True)
    acc_g = out["acc"].t().cpu().numpy()
    om = orc.make_model()
    _, acc_o, nrhs_o = orc.forward_batch(om, dcs, want_acc=True)
    for i, dc in enumerate(dcs):
        _check(acc_g[i], acc_o[i], dc)
    assert np.all(out["status"].cpu().numpy() == 0)
    assert np.all(out["filled"].cpu().numpy() == 500)
    # executed RHS count: the kernel skips the bit-identical restart evaluation of every interval
    # after the first, so it executes exactly (n_out - 2) fewer RHS than SciPy for the same steps
    nrhs_g = out["nrhs"].cpu().numpy()
    nonstiff = dcs >= 100
    assert np.array_equal(nrhs_g[nonstiff], nrhs_o[nonstiff] - 498)
    nstep_g = out["nstep"].cpu().numpy()
    assert np.all(nstep_g[nonstiff] == 503)
    # output times accumulate exactly like the reference (t_k+1 = t_k + delta_t inside the solver)
    tt = out["t"][:, 7].cpu().numpy()
    t_o, _, _ = orc.forward(orc.make_model(Dc=1000.0))
    assert np.array_equal(tt, t_o)


@pytest.mark.parametrize("damping", [True, False])
def test_sse_vs_oracle(cuda, pkg, orc, damping):
    g = load_golden("sse_grid.json")
    m = pkg.RateStateModel()
    m.RadiationDamping = damping
    out = m.evaluate_batch(g["grid"], data=g["data"], want_acc=False)
    sse_o, _, _ = orc.forward_batch(orc.make_model(radiation_damping=int(damping)), g["grid"], data=g["data"])
    assert np.allclose(out["sse"].cpu().numpy(), sse_o, rtol=1e-10, atol=0)
    if damping:
        assert np.allclose(out["sse"].cpu().numpy(), g["sse"], rtol=1e-10, atol=0)     # reference values


def test_silent_failure_is_flagged(cuda, pkg, orc):
    """Dc <= 1e-4: the reference bails in interval 1 with a zero tail (q9); the kernel flags it."""
    g = load_golden("forward_trajectories.json")
    case = [c for c in g["cases"] if "filled" in c][0]
    m = pkg.RateStateModel()
    out = m.evaluate_batch([case["Dc"], 1000.0, 5e-5], data=np.zeros(500))
    st = out["status"].cpu().numpy()
    assert st[0] == pkg._lib.CHAIN_NMAX and st[1] == 0 and st[2] == pkg._lib.CHAIN_NMAX
    assert out["filled"].cpu().numpy()[0] == case["filled"]
    acc = out["acc"][:, 0].cpu().numpy()
    assert np.all(acc[case["filled"]:] == 0.0) and acc[1] != 0.0
    # the partial value stored at k = 1 agrees with the oracle's (same failure point)
    _, acc_o, sto = orc.forward(orc.make_model(Dc=case["Dc"]))
    # (the failing step sequence is chaotic; agreement is to a few per cent, not to rounding)
    assert acc[1] == pytest.approx(acc_o[1], rel=0.05)
    m.Dc = case["Dc"]
    with pytest.warns(UserWarning, match="larger nsteps"):
        m.evaluate()


@pytest.mark.parametrize("c", [1, 31, 32, 33, 257, 5000])
def test_ragged_batch_sizes(cuda, pkg, orc, c):
    rng = np.random.default_rng(c)
    dcs = rng.uniform(200.0, 5000.0, size=c)
    data = orc.forward(orc.make_model(Dc=1350.0))[1]
    m = pkg.RateStateModel()
    out = m.evaluate_batch(dcs, data=data, want_acc=False)
    pick = rng.choice(c, size=min(c, 12), replace=False)
    sse_o, _, _ = orc.forward_batch(orc.make_model(), dcs[pick], data=data)
    assert np.allclose(out["sse"].cpu().numpy()[pick], sse_o, rtol=1e-9, atol=1e-18)


@pytest.mark.parametrize("n,t_end", [(200, 30.0), (501, 50.1), (1025, 102.5), (2100, 210.0)])
def test_other_grids_and_streamed_series(cuda, pkg, orc, n, t_end):
    """n > 1024 exercises the streaming (non-resident) TMA tile path, odd n the tail element."""
    dcs = np.array([300.0, 1350.0, 4000.0])
    om = orc.make_model(number_time_steps=n, end_time=t_end)
    assert orc.num_outputs(om) in (n, n - 1)
    data = orc.forward(orc.make_model(Dc=900.0, number_time_steps=n, end_time=t_end))[1]
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    assert m.num_outputs() == orc.num_outputs(om)
    out = m.evaluate_batch(dcs, data=data)
    sse_o, acc_o, _ = orc.forward_batch(om, dcs, data=data, want_acc=True)
    acc_g = out["acc"].t().cpu().numpy()
    for i, dc in enumerate(dcs):
        _check(acc_g[i], acc_o[i], dc)
    assert np.allclose(out["sse"].cpu().numpy(), sse_o, rtol=1e-9, atol=0)


def test_per_chain_a_b(cuda, pkg, orc):
    rng = np.random.default_rng(3)
    c = 40
    a = rng.uniform(0.008, 0.013, c)
    b = rng.uniform(0.012, 0.018, c)
    dc = rng.uniform(500, 3000, c)
    m = pkg.RateStateModel()
    out = m.evaluate_batch(dc, a=a, b=b)
    acc_g = out["acc"].t().cpu().numpy()
    for i in range(0, c, 7):
        _, acc_o, _ = orc.forward(orc.make_model(Dc=dc[i], a=a[i], b=b[i]))
        _check(acc_g[i], acc_o, dc[i])


def test_carry_mode_within_tolerance(cuda, pkg, orc):
    """CARRY (h and FSAL carried across output points) must stay inside the same trajectory gate."""
    dcs = np.array([60.0, 100.0, 1000.0, 1350.0, 5000.0])
    m = pkg.RateStateModel()
    m.integ_mode = "carry"
    out = m.evaluate_batch(dcs)
    acc_g = out["acc"].t().cpu().numpy()
    _, acc_o, _ = orc.forward_batch(orc.make_model(), dcs, want_acc=True)
    for i, dc in enumerate(dcs):
        _check(acc_g[i], acc_o[i], dc)
    assert np.all(out["nrhs"].cpu().numpy() < 6100)          # 12 instead of 13 executed RHS per interval


def test_vstep_loading_vs_oracle(cuda, pkg, orc):
    """VSTEP loading extension: same solver, piecewise-constant load-point velocity."""
    n, t_end = 600, 60.0
    kw = dict(loading=orc.LOAD_VSTEP, vstep_period=10.0, vstep_factor=3.0)
    dcs = np.array([2.0, 20.0, 200.0])
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    m.loading, m.vstep_period, m.vstep_factor = "vstep", 10.0, 3.0
    out = m.evaluate_batch(dcs)
    acc_g = out["acc"].t().cpu().numpy()
    _, acc_o, _ = orc.forward_batch(orc.make_model(number_time_steps=n, end_time=t_end, **kw), dcs, want_acc=True)
    assert np.all(out["status"].cpu().numpy() == 0)
    for i, dc in enumerate(dcs):
        scale = np.max(np.abs(acc_o[i]))
        assert np.max(np.abs(acc_g[i] - acc_o[i])) <= 1e-6 * scale


def test_size_independent_properties_at_scale(cuda, pkg):
    """Full-size batch (65,536 chains): identical parameters give identical results in every lane,
    SSE against the model's own trajectory is exactly zero, and SSE grows away from the truth."""
    m = pkg.RateStateModel()
    m.Dc = 1325.0
    _, acc, _ = m.evaluate()
    c = 65536
    dcs = np.full(c, 1325.0)
    dcs[1::2] = 2000.0
    out = m.evaluate_batch(dcs, data=acc, want_acc=False)
    sse = out["sse"].cpu().numpy()
    assert np.all(sse[0::2] == 0.0)
    assert np.all(sse[1::2] == sse[1]) and sse[1] > 0
    assert np.all(out["status"].cpu().numpy() == 0)


def test_unstable_step_is_rejected_after_velocity_jump(cuda, pkg, orc):
    """Stiff regime with a 10x load-velocity jump (cfg-4 style): right after the jump the first trial step is
    violently unstable; its error norm overflows and the step must be rejected like SciPy does (regression
    test for the division-free accept test, which compared inf <= inf)."""
    n, t_end = 2400, 240.0
    kw = dict(loading=orc.LOAD_VSTEP, vstep_period=120.0, vstep_factor=10.0)
    dcs = np.array([0.05, 0.1, 1.0])
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    m.loading, m.vstep_period, m.vstep_factor = "vstep", 120.0, 10.0
    out = m.evaluate_batch(dcs)
    assert np.all(out["status"].cpu().numpy() == 0)
    acc_g = out["acc"].t().cpu().numpy()
    for i in range(len(dcs)):
        assert np.all(np.isfinite(acc_g[i]))
        _check_stiff(orc, acc_g[i], dcs[i], f"vstep_jump_n2400_Dc{dcs[i]}", number_time_steps=n, end_time=t_end, **kw)


def test_batch_position_invariance_at_131072(cuda, pkg):
    """cfg-5 shard size: 131,072 parameter sets = 128 shuffled copies of 1,024 distinct Dc values.  A
    parameter set's SSE must not depend on its lane, warp or block, nor on the batch size."""
    m = pkg.RateStateModel()
    m.Dc = 1325.0
    np.random.seed(3)
    _, _, data = m.evaluate()
    rng = np.random.default_rng(11)
    base = rng.uniform(200.0, 5000.0, 1024)
    idx = rng.permutation(np.repeat(np.arange(1024), 128))
    big = m.evaluate_batch(base[idx], data=data)
    small = m.evaluate_batch(base, data=data)
    sse_small = small["sse"].cpu().numpy()
    assert np.array_equal(big["sse"].cpu().numpy(), sse_small[idx])
    assert np.array_equal(big["nrhs"].cpu().numpy(), small["nrhs"].cpu().numpy()[idx])
    assert np.all(big["status"].cpu().numpy() == 0)


def test_long_series_is_causal_and_streamed_sse_is_consistent(cuda, pkg):
    """cfg-4 grid (N = 1e5 output points, velocity steps every 1,000 s) in the non-stiff regime: every
    output interval restarts the integrator, so the first 1,000 points of the long solve equal the
    solve that stops there, bit for bit; the SSE accumulated from the TMA-streamed series equals the
    sum over the returned trajectory; the last output time is the accumulated grid of the reference."""
    n = 100_000
    kw = dict(loading="vstep", vstep_period=1000.0, vstep_factor=10.0)
    long_m = pkg.RateStateModel(number_time_steps=n, end_time=n * 0.1)
    short_m = pkg.RateStateModel(number_time_steps=1000, end_time=100.0)
    for m in (long_m, short_m):
        m.loading, m.vstep_period, m.vstep_factor = kw["loading"], kw["vstep_period"], kw["vstep_factor"]
    dcs = np.array([400.0, 1000.0, 1325.0, 2500.0, 6000.0])
    rng = np.random.default_rng(5)
    data = rng.standard_normal(n)
    lo = long_m.evaluate_batch(dcs, data=data, want_acc=True, want_t=True)
    sh = short_m.evaluate_batch(dcs, want_acc=True, want_t=True)
    acc_l, acc_s = lo["acc"].cpu().numpy(), sh["acc"].cpu().numpy()        # [n_out, C]
    assert acc_l.shape == (n, 5) and acc_s.shape == (1000, 5)
    assert np.array_equal(acc_l[:1000], acc_s)
    assert np.array_equal(lo["t"].cpu().numpy()[:1000], sh["t"].cpu().numpy())
    assert np.all(lo["status"].cpu().numpy() == 0) and np.all(lo["filled"].cpu().numpy() == n)
    # the sampler's SSE is a plain running sum over k (MCMC.py:387 semantics, sequential order)
    ref = np.zeros(5)
    for k in range(n):
        e = acc_l[k] - data[k]
        ref += e * e
    assert np.allclose(lo["sse"].cpu().numpy(), ref, rtol=1e-12, atol=0.0)
    # output times are accumulated, t_{k+1} = t_k + delta_t (RateStateModel.py:382-384)
    t = 0.0
    for _ in range(n - 1):
        t += long_m.delta_t
    assert lo["t"].cpu().numpy()[-1, 0] == t
    # the velocity steps are felt: acceleration spikes right after t = 1000 s
    assert np.max(np.abs(acc_l[10_000:10_020, 2])) > 100 * np.max(np.abs(acc_l[9_900:9_990, 2]))


def test_stiff_variant_on_reference_loading_and_both_variants_on_vstep(cuda, pkg, orc):
    """The kernels exist in two variants (rsfm_kernels.cu: stiff_variant): the default one and the one used for
    velocity-step loading, which re-bases the friction law on the current load level, resumes the general-range
    step from the first stage that left the fast ranges and uses the SFU-seeded controller root.  Forced onto the
    reference's own loading (cfg.solver_variant = RSFM_VARIANT_STIFF) it must stay inside the golden-trajectory gate; on a velocity-step
    problem both variants must agree with the oracle and with each other."""
    for case in load_golden("forward_trajectories.json")["cases"]:
        if "filled" in case or not case["RadiationDamping"]:
            continue
        m = pkg.RateStateModel(number_time_steps=case["N"], end_time=case.get("end_time", 50.0))
        m.solver_variant = "stiff"
        m.Dc = case["Dc"]
        _, acc, _ = m.evaluate()
        _check(acc, case["acc"], case["Dc"])
    # velocity steps, stiff and non-stiff chains in one batch
    n, t_end = 1200, 120.0
    kw = dict(loading=orc.LOAD_VSTEP, vstep_period=30.0, vstep_factor=10.0)
    dcs = np.array([0.05, 0.3, 2.0, 40.0, 400.0])
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    m.loading, m.vstep_period, m.vstep_factor = "vstep", 30.0, 10.0
    floors = [oracle_ulp_floor(orc, dc, number_time_steps=n, end_time=t_end, **kw) for dc in dcs]
    # "1x": the stiff variant scoring every step that left the fast ranges with the general-range stages
    # (cfg.stiff_exact) instead of taking the exploding trial steps as rejected
    res = {}
    for v, exact in (("0", "0"), ("1", "0"), ("1", "1")):
        m.solver_variant = "stiff" if v == "1" else "default"
        m.stiff_exact = exact == "1"
        o = m.evaluate_batch(dcs)
        assert np.all(o["status"].cpu().numpy() == 0)
        res[v + ("x" if exact == "1" else "")] = (o["acc"].t().cpu().numpy(), o["nstep"].cpu().numpy())
    for i in range(len(dcs)):
        acc_o, scale, floor = floors[i]
        gate = max(RTOL_STIFF if dcs[i] < 50 else RTOL, STIFF_FLOOR_MULT * floor)
        errs = {v: float(np.max(np.abs(res[v][0][i] - acc_o))) / scale for v in res}
        _record(f"vstep_variants_n1200_Dc{dcs[i]}", Dc=dcs[i], oracle_ulp_floor_rel=floor, gate_rel=gate,
                **{"err_rel_" + v: e for v, e in errs.items()})
        for v in res:
            assert errs[v] <= gate, (v, dcs[i], errs[v], gate)
        assert np.max(np.abs(res["0"][0][i] - res["1"][0][i])) <= 2 * gate * scale
    # same controller semantics: attempted-step counts of the variants differ by about 1 % at most in the stiff
    # regime (observed 0.2-1.2 %: the step sequences of two roundings of the same arithmetic part there) and not at
    # all outside it
    for v in ("1", "1x"):
        assert np.all(np.abs(res["0"][1].astype(float) - res[v][1]) <= 0.02 * res["0"][1]), v
        assert np.array_equal(res["0"][1][dcs >= 40.0], res[v][1][dcs >= 40.0]), v


def test_stiff_rule_changes_no_decision(cuda, pkg):
    """The stiff variant takes a trial step as rejected without scoring it when the load is constant at the frame's
    base level over the step, the step starts inside half the fast ranges and a stage leaves them.
    cfg.stiff_exact scores every such step with the general-range stages.  If the rule never overrules the exact
    arithmetic the two runs take the same steps and return the SAME BITS.  (Round 2 found the counter-example
    that shaped the load condition: the accumulated output times put a velocity jump a few ulp inside an interval
    whose frame is the old level; the first step after the jump starts at the old steady state, legitimately
    leaves the ranges and is accepted -- profiles/microbench/stiff_wild_probe.py.)"""
    for n, t_end, period, factor in ((1200, 120.0, 30.0, 10.0), (600, 60.0, 20.0, 3.0)):
        dcs = np.array([0.03, 0.05, 0.08, 0.3, 0.6, 1.0, 1.5, 2.0, 3.0, 5.0, 10.0, 40.0])
        m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
        m.loading, m.vstep_period, m.vstep_factor = "vstep", period, factor
        m.solver_variant = "stiff"
        out = {}
        for exact in (False, True):
            m.stiff_exact = exact
            o = m.evaluate_batch(dcs)
            assert np.all(o["status"].cpu().numpy() == 0)
            out[exact] = (o["acc"].cpu().numpy(), o["nstep"].cpu().numpy())
        assert np.array_equal(out[False][1], out[True][1]), (period, out[False][1], out[True][1])
        assert np.array_equal(out[False][0], out[True][0]), period


def test_stiff_variant_per_chain_a_b(cuda, pkg, orc):
    """Velocity-step loading with per-chain (a, b, Dc) -- the forward solves of a joint-posterior sampler in the stiff
    regime: the re-based reference friction mu_ref + (a - b) ln(lambda) differs from chain to chain."""
    rng = np.random.default_rng(5)
    c, n, t_end = 12, 600, 60.0
    a = rng.uniform(0.008, 0.013, c)
    b = rng.uniform(0.012, 0.018, c)
    dc = np.concatenate([rng.uniform(0.04, 0.3, c // 2), rng.uniform(1.0, 300.0, c - c // 2)])
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    m.loading, m.vstep_period, m.vstep_factor = "vstep", 15.0, 10.0
    out = m.evaluate_batch(dc, a=a, b=b)
    assert np.all(out["status"].cpu().numpy() == 0)
    acc_g = out["acc"].t().cpu().numpy()
    for i in range(c):
        _check_stiff(orc, acc_g[i], dc[i], f"vstep_abdc_n600_chain{i}", a=a[i], b=b[i], number_time_steps=n, end_time=t_end,
                     loading=orc.LOAD_VSTEP, vstep_period=15.0, vstep_factor=10.0)


@pytest.mark.parametrize("dcs", [(0.03, 0.05, 0.08)])
def test_cfg4_full_size_trajectories_vs_oracle(cuda, pkg, orc, dcs):
    """cfg 4 at its stated size: N = 100,000 output points over 10,000 s, load velocity x10 every
    1,000 s, Dc in the stability-limited regime -- single trajectories on the GPU (stiff kernel variant, TMA-streamed
    series) against the C oracle (about 2e7 RHS evaluations each).  Gate and recorded numbers as in _check_stiff;
    the SSE accumulated against a streamed series equals the oracle's sum over ITS trajectory to the same level."""
    n, t_end = 100_000, 10_000.0
    kw = dict(number_time_steps=n, end_time=t_end, loading=orc.LOAD_VSTEP, vstep_period=1000.0, vstep_factor=10.0)
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    m.loading, m.vstep_period, m.vstep_factor = "vstep", 1000.0, 10.0
    rng = np.random.default_rng(4)
    truth = orc.forward(orc.make_model(Dc=0.05, **kw))[1]
    # 2 % multiplicative noise: with 20 % the noise cross term (+-2.2) exceeds sum (acc(Dc) - truth)^2 (0.16, 0.96)
    # and the SSE -- the oracle's own just the same -- no longer orders the three Dc values
    data = truth + 0.02 * np.abs(truth) * rng.standard_normal(n)
    out = m.evaluate_batch(np.array(dcs), data=data, want_t=True)
    assert np.all(out["status"].cpu().numpy() == 0) and np.all(out["filled"].cpu().numpy() == n)
    acc_g = out["acc"].t().cpu().numpy()
    sse_g = out["sse"].cpu().numpy()
    nrhs_g = out["nrhs"].cpu().numpy()
    for i, dc in enumerate(dcs):
        err, floor = _check_stiff(orc, acc_g[i], dc, f"cfg4_full_N100000_Dc{dc}", **kw)
        acc_o = orc.forward(orc.make_model(Dc=dc, **kw))[1]
        sse_o = float(np.sum((acc_o - data) ** 2))
        sse_self = float(np.sum((acc_g[i] - data) ** 2))
        assert sse_g[i] == pytest.approx(sse_self, rel=1e-11)                 # streamed SSE = sum over its own trajectory
        assert sse_g[i] == pytest.approx(sse_o, rel=max(1e-5, 50 * floor))
        assert 1.0e7 < nrhs_g[i] < 6.0e7                                      # ~2e7 RHS per solve
    # the likelihood the sampler sees is usable at this size: SSE is smallest at the true Dc
    assert np.argmin(sse_g) == 1


def _load_table(n, dt):
    tt = np.arange(n + 1) * dt
    return 0.5 * np.sin(0.7 * tt) * np.exp(-tt / 30.0) + 0.3 * (tt > 20.0)


@pytest.mark.parametrize("law, loading", [("slip", "sine_decay"), ("aging", "table"), ("slip", "table")])
def test_slip_law_and_tabulated_loading_vs_oracle(cuda, pkg, orc, law, loading):
    """SURVEY 8f.4 extensions on the device: Ruina's slip law (general-range stages only) and a piecewise-linear
    tabulated load, against the C oracle (itself bit-identical to SciPy with the one line swapped,
    tests/test_oracle_vs_scipy.py): trajectory gate 1e-9 outside the stiff regime, SSE and step counters too."""
    tab = _load_table(500, 0.1)
    dcs = np.array([120.0, 300.0, 1350.0, 5000.0, 9000.0])
    for damping in (True, False):
        m = pkg.RateStateModel()
        m.RadiationDamping = damping
        m.state_law, m.loading = law, loading
        m.load_table, m.load_dt = tab, 0.1
        rng = np.random.default_rng(7)
        kw = dict(radiation_damping=int(damping), state_law=orc.LAW_SLIP if law == "slip" else orc.LAW_AGING)
        if loading == "table":
            kw.update(loading=orc.LOAD_TABLE, load_table=tab, load_dt=0.1)
        truth = orc.forward(orc.make_model(Dc=1350.0, **kw))[1]
        data = truth + np.abs(truth) * rng.standard_normal(truth.size)
        out = m.evaluate_batch(dcs, data=data)
        assert np.all(out["status"].cpu().numpy() == 0)
        acc_g = out["acc"].t().cpu().numpy()
        for i, dc in enumerate(dcs):
            t_o, acc_o, st = orc.forward(orc.make_model(Dc=dc, **kw))
            err = _check(acc_g[i], acc_o, dc)
            _record(f"ext_{law}_{loading}_damp{int(damping)}_Dc{dc}", Dc=dc, err_rel=err)
            assert int(out["nstep"][i]) == st.nstep
            assert out["sse"][i].item() == pytest.approx(float(np.sum((acc_o - data) ** 2)), rel=1e-9)
    # evaluate() through the model protocol
    m.Dc = 1350.0
    _, acc1, _ = m.evaluate()
    assert np.array_equal(acc1, acc_g[2])


def test_slip_law_in_the_stiff_regime_and_rhs(cuda, pkg, orc):
    """The slip law under velocity steps (stiff variant; gate = measured floor as for the aging law) and the RHS
    hook: rsfm_rhs_eval(general = 1) evaluates theta' = -(v theta/Dc) ln(v theta/Dc) like the oracle."""
    import ctypes as C
    torch = cuda
    n, t_end = 600, 60.0
    kw = dict(number_time_steps=n, end_time=t_end, loading=orc.LOAD_VSTEP, vstep_period=15.0, vstep_factor=10.0,
              state_law=orc.LAW_SLIP)
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    m.loading, m.vstep_period, m.vstep_factor, m.state_law = "vstep", 15.0, 10.0, "slip"
    dcs = np.array([0.08, 0.5, 30.0, 400.0])
    out = m.evaluate_batch(dcs)
    assert np.all(out["status"].cpu().numpy() == 0)
    acc_g = out["acc"].t().cpu().numpy()
    for i, dc in enumerate(dcs):
        _check_stiff(orc, acc_g[i], dc, f"vstep_slip_n600_Dc{dc}", **kw)
    # RHS at states away from steady state
    rng = np.random.default_rng(2)
    k = 64
    dc = rng.uniform(50.0, 5000.0, k)
    th = dc * rng.uniform(0.7, 1.4, k)
    mu = 0.6 + rng.uniform(-0.01, 0.01, k)
    t = rng.uniform(0.0, 50.0, k)
    m2 = pkg.RateStateModel()
    m2.state_law = "slip"
    cfg = m2.to_cfg()
    args = [torch.from_numpy(x).cuda() for x in (t, mu, th, dc)]
    got = torch.empty((3, k), dtype=torch.float64, device="cuda")
    pkg._lib.check(pkg._lib.load().rsfm_rhs_eval(C.byref(cfg), k, *[a.data_ptr() for a in args], None, None, 1,
                                                 got.data_ptr(), None))
    got = got.cpu().numpy().T
    for i in range(k):
        want = orc.rhs(orc.make_model(Dc=dc[i], state_law=orc.LAW_SLIP), t[i], [mu[i], th[i], 1.0])
        assert np.allclose(got[i], want, rtol=1e-12, atol=1e-15 * np.abs(want).max()), i
