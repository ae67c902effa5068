"""k1 (radiation-damping coefficient, RateStateModel.py:171, 349-353) as the sampled / batched constant -- SURVEY 8f.4
"named parameters (a, b, Dc, k1 ...)".  The CUDA path (RSFM_PARAM_K1) against the CPU oracle, which is pinned on
trajectories of the unmodified reference with ``model.k1`` varied (tests/golden/forward_k1.json)."""
import numpy as np
import pytest

from conftest import load_golden
from test_gpu_mcmc import _run_with_draws

pytestmark = pytest.mark.gpu

TRAJ_GATE = 1e-9          # SURVEY 8c, non-stiff regime


def test_forward_batch_over_k1_matches_reference_trajectories(cuda, pkg, orc):
    """Batch axis = k1 (values on both sides of the library's switch between its fast and general-range stages at
    4e-7, zero included), Dc common: every trajectory against the golden vectors of the unmodified reference where one
    exists and against the oracle everywhere; SSE(k1) against the reference's grid."""
    g = load_golden("forward_k1.json")
    for dc in (100.0, 1000.0, 1350.0):
        cases = [c for c in g["cases"] if c["Dc"] == dc]
        k1 = np.array([c["k1"] for c in cases] + [3.9e-7, 4.1e-7, 2e-6, 5e-4, 9.9e-3])
        model = pkg.RateStateModel()
        out = model.evaluate_batch(dc, k1=k1, want_t=True)
        acc = out["acc"].t().cpu().numpy()
        assert int(out["status"].abs().sum().item()) == 0
        worst = 0.0
        for i, k in enumerate(k1):
            ref = cases[i]["acc"] if i < len(cases) else orc.forward(orc.make_model(Dc=dc, k1=float(k)))[1]
            worst = max(worst, float(np.max(np.abs(acc[i] - ref)) / np.max(np.abs(ref))))
        assert worst < TRAJ_GATE, (dc, worst)
        # the same step sequence as the reference's solver: RHS count of the oracle
        # (the kernel skips the bit-identical restart evaluation of every interval after the first: n_out - 2 fewer)
        for i in (0, 3, len(k1) - 1):
            st = orc.forward(orc.make_model(Dc=dc, k1=float(k1[i])))[2]
            assert int(out["nstep"][i].item()) == st.nstep
            assert int(out["nrhs"][i].item()) == st.nrhs - 498
    model = pkg.RateStateModel()
    out = model.evaluate_batch(g["data_Dc"], k1=g["k1_grid"], data=g["data"], want_acc=False)
    assert np.allclose(out["sse"].cpu().numpy(), g["sse"], rtol=1e-9, atol=0)


def test_k1_batch_needs_one_dc(cuda, pkg):
    model = pkg.RateStateModel()
    with pytest.raises(ValueError):
        model.evaluate_batch([1000.0, 1100.0], k1=[1e-3, 2e-3])


def test_k1_cfg_validation(built_lib, pkg):
    import ctypes as C
    lib = built_lib
    cfg = pkg._lib.default_cfg()
    cfg.sampled_param = pkg._lib.PARAM_K1
    cfg.dc_fixed = 0.0
    assert lib.rsfm_create(C.byref(cfg), 4, 1, 0) is None
    assert b"dc_fixed" in lib.rsfm_last_error()
    cfg.dc_fixed, cfg.n_params = 1000.0, 3
    assert lib.rsfm_create(C.byref(cfg), 4, 1, 0) is None
    cfg.n_params, cfg.sampled_param = 1, 7
    assert lib.rsfm_create(C.byref(cfg), 4, 1, 0) is None


def test_deterministic_k1_chain_replays_on_the_oracle(cuda, pkg, orc):
    """Host-supplied absolute proposals, uniforms and gammas (list priors: no adaptation, q2): every decision, every
    sample bit for bit, sigma^2 to 1e-8; start covariance from the forward difference in k1 (MCMC.py:245-266 with
    model.k1 in the place of model.Dc)."""
    g = load_golden("forward_k1.json")
    data = g["data"]
    rng = np.random.default_rng(5)
    ns, q0, lo, hi = 48, 2e-3, 1.8e-3, 3.6e-3
    prop = 2.7e-3 + 5e-4 * rng.standard_normal(ns)
    u, gam = rng.random(ns), rng.gamma(0.5 * (0.01 + 500), size=ns)
    model = pkg.RateStateModel()
    model.Dc = 1000.0
    mc = pkg.MCMC(model, data, 1000.0, ["Uniform", lo, hi], q0, nsamples=ns, verbose=False, param_names=("k1",),
                  deterministic_inputs={"proposals": prop, "uniforms": u, "gammas": gam})
    out = mc.sample(False)
    om = orc.make_model(Dc=1000.0, sampled_param=orc.PARAM_K1)
    chain_o, s2_o, acc_o, vstart, _ = orc.chain_replay(om, data, q0, lo, hi, 3, ns, prop, u, gam)
    assert mc.Vstart[0, 0] == pytest.approx(vstart, rel=1e-5)          # limit of the 1e-6 forward difference
    assert np.array_equal(mc.accepts, acc_o)
    assert 0.1 < acc_o.mean() < 0.9 and np.any((prop <= lo) | (prop >= hi))
    assert np.array_equal(out[0], chain_o[mc.nburn:])
    assert np.allclose(mc.std2, s2_o[mc.nburn:], rtol=1e-8, atol=0)
    assert model.Dc == 1000.0 and 0.0 < model.k1 < 0.01                # MCMC mutates the sampled attribute (q6)


def test_k1_chain_with_the_references_adaptation_on_the_device(cuda, pkg, orc):
    """Dict priors, standard-normal draws in: proposals q + sqrt(V) z with V = Vstart and, after the first boundary,
    the reference's windowed update (MCMC.py:200-204) formed on the device.  Vstart comes from a 1e-6 forward
    difference of two solves that agree with the oracle's to 1e-12, i.e. it carries ~1e-6 of relative noise on either
    side, so the samples are compared to that fraction of a proposal step; the decisions must all agree.  The run
    stops before a second boundary: the update uses the Cholesky FACTOR as a covariance (q3), which for a parameter of
    size 1e-3 gives proposals of s.d. ~1e-2, nearly everything leaves the prior box, and a window of ten equal
    samples then decides on the last bit of its mean whether the next scale is unchanged or ~1e-19 -- in the
    reference as much as here."""
    g = load_golden("forward_k1.json")
    data = g["data"]
    rng = np.random.default_rng(5)
    ns, q0 = 18, 2e-3
    z, u, gam = rng.standard_normal(ns), rng.random(ns), rng.gamma(0.5 * (0.01 + 500), size=ns)
    model = pkg.RateStateModel()
    model.Dc = 1000.0
    mc = pkg.MCMC(model, data, 1000.0, {1: 0.0, 2: 0.01}, q0, nsamples=ns, verbose=False, param_names=("k1",),
                  deterministic_inputs={"z": z, "uniforms": u, "gammas": gam})
    out = mc.sample(False)
    om = orc.make_model(Dc=1000.0, sampled_param=orc.PARAM_K1)
    chain_o, s2_o, acc_o, vstart, _ = orc.chain_replay(om, data, q0, 0.0, 0.01, 2, ns, z, u, gam, compat_adapt=True)
    assert mc.Vstart[0, 0] == pytest.approx(vstart, rel=1e-5)
    assert np.array_equal(mc.accepts, acc_o)
    assert acc_o[:10].mean() > 0.3
    err = float(np.max(np.abs(out[0] - chain_o[mc.nburn:])))
    assert err < 1e-5 * np.sqrt(vstart) * ns, err
    assert np.allclose(mc.std2, s2_o[mc.nburn:], rtol=1e-6, atol=0)


def test_free_running_k1_chains_replay_through_oracle(cuda, pkg, orc):
    """Philox-driven chains over k1 (bounds tight enough for out-of-bounds proposals), replayed with the draws the
    kernel reports; chains start on both sides of the fast / general switch."""
    torch = cuda
    g = load_golden("forward_k1.json")
    data = g["data"]
    model = pkg.RateStateModel()
    cfg = model.to_cfg()
    cfg.n_params, cfg.n_prior_len = 1, 3
    cfg.sampled_param, cfg.dc_fixed = pkg._lib.PARAM_K1, 1000.0
    cfg.lo[0], cfg.hi[0] = 2.0e-3, 3.6e-3
    c, ns = 40, 30
    samples, s2, acc, draws, depth, tot = _run_with_draws(torch, pkg, cfg, c, 77, 300, [2.8e-3], data, [ns])
    assert depth == 0                                                  # one-thread-per-chain kernel
    assert 0 < acc.mean() < 1
    om = orc.make_model(Dc=1000.0, sampled_param=orc.PARAM_K1)
    n_oob = 0
    for ch in (0, 9, 31, 39):
        u = draws[:, 1, ch]
        n_oob += int(np.isnan(u).sum())
        chain_o, s2_o, acc_o, _, _ = orc.chain_replay(om, data, 2.8e-3, 2.0e-3, 3.6e-3, 3, ns, draws[:, 0, ch],
                                                      np.nan_to_num(u, nan=0.5), draws[:, 2, ch])
        assert np.array_equal(acc[:, ch], acc_o), ch
        assert np.array_equal(samples[:, 0, ch], chain_o[1:]), ch
        assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-8, atol=0)
    assert n_oob > 0


def test_k1_posterior_recovers_the_truth(cuda, pkg):
    """4,096 chains, pooled adaptive Metropolis over k1: posterior mean within the posterior width of the value the
    data were generated with (3e-3), and the grid minimum of the reference's SSE(k1) inside it."""
    g = load_golden("forward_k1.json")
    model = pkg.RateStateModel()
    model.Dc = 1000.0
    rng = np.random.default_rng(3)
    c = 4096
    q0 = rng.uniform(1.5e-3, 4.5e-3, c)
    mc = pkg.MCMC(model, g["data"], 1000.0, ["Uniform", 0.0, 0.01], q0, nsamples=300, verbose=False, n_chains=c,
                  seed=9, param_names=("k1",), adapt="pooled", adapt_start=50)
    out = mc.sample(False)                                             # [C, 1, n]
    assert mc.stats["failed_chains"] == 0
    mean, sd = float(out.mean()), float(out.std())
    assert 0 < sd < 1e-3
    assert abs(mean - 3e-3) < 3 * sd
    best = g["k1_grid"][int(np.argmin(g["sse"]))]
    assert abs(mean - best) < 3 * sd
    assert 0.05 < float(np.mean(mc.acceptance_ratio)) < 0.9
