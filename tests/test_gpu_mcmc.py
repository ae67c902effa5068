"""Parity of the fused CUDA MCMC kernel with the reference chain (golden, step for step), the CPU
oracle (free-running draws replayed), and the exact posterior (quadrature)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _priors(g):
    return {1: g["lo"], 2: g["hi"]} if g["qpriors_form"] == "dict" else ["Uniform", g["lo"], g["hi"]]


@pytest.mark.parametrize("name", ["chain_list_priors.json", "chain_dict_priors.json", "chain_bounds.json", "chain_cfg1_full.json"])
def test_deterministic_replay_of_reference_chain(cuda, pkg, name):
    """Host-supplied proposals, uniforms and gamma draws recorded from the UNMODIFIED reference:
    the kernel must take the same accept/reject decision at every step (north_star check 2)."""
    g = load_golden(name)
    ns, nb = g["nsamples"], g["nburn"]
    model = pkg.RateStateModel()
    mc = pkg.MCMC(model, g["data"], g["dc_true"], _priors(g), g["qstart"], nsamples=ns, verbose=False,
                  compat_adapt=False,
                  deterministic_inputs={"proposals": g["proposals"], "uniforms": g["uniforms"],
                                        "gammas": g["gammas_unit"]})
    out = mc.sample(False)
    assert out.shape == (1, ns + 1 - nb) and out.dtype == np.float64          # MCMC.py:544
    assert np.array_equal(mc.accepts, g["accepts"])
    assert np.array_equal(out, g["chain_post_burn"])                            # accepted values are the inputs
    assert mc.std2.shape == (ns + 1 - nb,)
    assert np.allclose(mc.std2, g["std2_post_burn"], rtol=1e-8, atol=0)
    assert mc.Vstart.shape == (1, 1)
    assert mc.Vstart.item() == pytest.approx(g["Vstart"].item(), rel=1e-6)
    assert mc.nburn == nb and mc.stats["failed_chains"] == 0


def test_compat_adaptation_on_device(cuda, pkg, orc):
    """dict priors: standard-normal draws in, the kernel applies the reference's (quirky) adaptation
    itself and must reproduce the proposal scales, hence the whole chain."""
    g = load_golden("chain_dict_priors.json")
    ns, nb = g["nsamples"], g["nburn"]
    uni = np.nan_to_num(g["uniforms"], nan=0.5)
    full, _, _, _, _ = orc.chain_replay(orc.make_model(), g["data"], g["qstart"], g["lo"], g["hi"],
                                        g["n_prior_len"], ns, g["proposals"], uni, g["gammas_unit"])
    z = (g["proposals"] - full[:-1]) / np.sqrt(g["V_used"])
    mc = pkg.MCMC(pkg.RateStateModel(), g["data"], g["dc_true"], _priors(g), g["qstart"], nsamples=ns,
                  verbose=False, deterministic_inputs={"z": z, "uniforms": g["uniforms"], "gammas": g["gammas_unit"]})
    assert mc.compat_adapt
    out = mc.sample(False)
    assert np.array_equal(mc.accepts, g["accepts"])
    assert np.allclose(out, g["chain_post_burn"], rtol=1e-9, atol=0)


def test_free_running_chains_replay_through_oracle(cuda, pkg, orc):
    """Philox-driven chains: dump the draws the kernel used, replay them on the CPU oracle."""
    import ctypes as C
    torch = cuda
    g = load_golden("sse_grid.json")
    lib = pkg._lib.load()
    model = pkg.RateStateModel()
    cfg = model.to_cfg()
    cfg.n_params, cfg.n_prior_len = 1, 3
    cfg.lo[0], cfg.hi[0] = 1270.0, 1380.0                # tight bounds: out-of-bounds proposals do occur
    c, ns = 48, 40
    dev = torch.device("cuda", 0)
    q0 = torch.full((1, c), 1300.0, dtype=torch.float64, device=dev)
    data_t = torch.from_numpy(g["data"]).to(dev)
    h = lib.rsfm_create(C.byref(cfg), c, 1234, 1000)
    assert h
    try:
        pkg._lib.check(lib.rsfm_init(h, q0.data_ptr(), data_t.data_ptr(), None))
        samples = torch.empty((ns, 1, c), dtype=torch.float64, device=dev)
        s2 = torch.empty((ns, c), dtype=torch.float64, device=dev)
        acc = torch.empty((ns, c), dtype=torch.uint8, device=dev)
        draws = torch.empty((ns, 3, c), dtype=torch.float64, device=dev)
        pkg._lib.check(lib.rsfm_run(h, ns, samples.data_ptr(), s2.data_ptr(), acc.data_ptr(), draws.data_ptr(), None))
        torch.cuda.synchronize()
    finally:
        lib.rsfm_destroy(h)
    samples, s2, acc, draws = (x.cpu().numpy() for x in (samples, s2, acc, draws))
    n_oob = 0
    for ch in (0, 1, 17, 47):
        prop, u, gam = draws[:, 0, ch], draws[:, 1, ch], draws[:, 2, ch]
        n_oob += int(np.isnan(u).sum())
        assert np.all((u[~np.isnan(u)] > 0) & (u[~np.isnan(u)] < 1)) and np.all(gam > 0)
        chain_o, s2_o, acc_o, _, _ = orc.chain_replay(orc.make_model(), g["data"], 1300.0, 1270.0, 1380.0, 3, ns,
                                                      prop, np.nan_to_num(u, nan=0.5), gam)
        assert np.array_equal(acc[:, ch], acc_o)
        assert np.array_equal(samples[:, 0, ch], chain_o[1:])
        assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-8, atol=0)
    assert n_oob > 0
    # different chains use different streams
    assert len(np.unique(draws[0, 0, :])) == c


def test_sharding_invariance(cuda, pkg):
    """Philox is keyed by the GLOBAL chain id: 64 chains in one sampler == two samplers of 32."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    kw = dict(nsamples=12, verbose=False, seed=77)
    full = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], 1000.0, n_chains=64, **kw).sample(False)
    assert full.shape == (64, 1, 7)

    import importlib
    sh = importlib.import_module("bayesian-markov-chain-monte-carlo_b200.sharding")
    parts = []
    for rank in (0, 1):
        mc = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], 1000.0, n_chains=64, shard=True, **kw)
        orig = sh.ChainShard.for_current_rank
        sh.ChainShard.for_current_rank = staticmethod(lambda total, r=rank: sh.ChainShard.for_rank(total, r, 2))
        try:
            parts.append(mc.sample(False))
        finally:
            sh.ChainShard.for_current_rank = orig
    assert np.array_equal(np.concatenate(parts, axis=0), full)
    assert not np.array_equal(full[0], full[1])


def test_posterior_moments_match_quadrature(cuda, pkg):
    """North-star check 3: posterior mean / sd of many GPU chains agree with the exact posterior.

    With sigma^2 integrated out under its conjugate update the marginal posterior of Dc is
    p(Dc) ~ (n0*s0 + SSE(Dc))^(-(n0+N)/2) only approximately (the reference's Gibbs step conditions
    on the previous sigma^2 through n0 = 0.01, negligible against N = 500), so the gate is MCSE-based
    with a small model-error allowance."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    grid = np.linspace(900.0, 1900.0, 4001)
    sse = model.evaluate_batch(grid, data=g["data"], want_acc=False)["sse"].cpu().numpy()
    n = 500
    logp = -0.5 * n * np.log(sse)
    p = np.exp(logp - logp.max())
    p /= p.sum()
    mean_q = float((grid * p).sum())
    sd_q = float(np.sqrt(((grid - mean_q) ** 2 * p).sum()))

    mc = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], 1000.0, nsamples=600, n_chains=512,
                  verbose=False, seed=5)
    out = mc.sample(False)                       # [512, 1, 301]
    draws = out[:, 0, :]
    chain_means = draws.mean(axis=1)
    mean_g = float(chain_means.mean())
    mcse = float(chain_means.std(ddof=1) / np.sqrt(len(chain_means)))
    sd_g = float(draws.std())
    assert abs(mean_g - mean_q) < 5 * mcse + 0.002 * sd_q, (mean_g, mean_q, mcse)
    assert abs(sd_g / sd_q - 1) < 0.05, (sd_g, sd_q)
    assert 0.2 < mc.acceptance_ratio.mean() < 0.95
    d = mc.diagnostics()
    assert d["rhat"][0] < 1.1 and d["ess"][0] > 512 * 5


def test_reference_api_surface(cuda, pkg, capsys):
    """Drop-in behaviours: printing, model mutation (q6), attribute names, JSON sample output."""
    import os
    import tempfile
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel(number_time_steps=500)
    np.random.seed(2024)
    mc = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 10000.0], 1000.0, nsamples=10)
    out = mc.sample(False)
    printed = capsys.readouterr().out.splitlines()
    assert printed[0] in ("0 True", "0 False") and printed[1].startswith("Generated Sample ----  ")
    assert printed[-1].startswith("acceptance ratio: ")
    assert out.shape == (1, 6) and mc.std2.shape == (6,)
    assert isinstance(model.Dc, np.ndarray) and model.Dc.shape == (1,)
    for attr in ("model", "qstart", "qpriors", "nsamples", "nburn", "verbose", "adapt_interval", "data",
                 "lstm_model", "n0", "qstart_limits", "dc_true", "std2", "Vstart"):
        assert hasattr(mc, attr)
    # np.random.seed makes the run reproducible, like the reference's global-RNG behaviour
    np.random.seed(2024)
    out2 = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 10000.0], 1000.0, nsamples=10, verbose=False).sample(False)
    assert np.array_equal(out, out2)
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "samples.json")
        mc.save_samples(fn)
        back = pkg.load_object(fn)
        assert np.array_equal(back["samples"][0], out)
        assert open(fn).read().count('"__ndarray__": true') == 3


@pytest.mark.parametrize("d,names,q0,bounds,priors", [
    (1, ("Dc",), 1000.0, None, ["Uniform", 0.0, 1e4]),
    (3, ("a", "b", "Dc"), np.array([0.0105, 0.0145, 1200.0]), [[0.005, 0.02], [0.005, 0.03], [0.0, 10000.0]],
     ["Uniform", 0.0, 1e4]),
    (1, ("Dc",), 1000.0, None, {1: 0.0, 2: 1e4}),         # dict priors: the reference's adaptation every 10 samples
    # a tight box: most proposals fall outside it, so the sequential d = 3 kernel walks lanes through
    # different numbers of solve-free iterations between solves
    (3, ("a", "b", "Dc"), np.array([0.0105, 0.0145, 1200.0]), [[0.0103, 0.0107], [0.0143, 0.0147], [1150.0, 1250.0]],
     ["Uniform", 0.0, 1e4]),
])
def test_speculative_kernel_is_bit_identical(cuda, pkg, d, names, q0, bounds, priors):
    """Speculative (prefetching) Metropolis evaluates the tree of the next g iterations in parallel; it
    must return exactly the chains of the sequential kernel: same samples, sigma^2, accept flags --
    also when the reference's windowed adaptation changes the proposal scale on the way (rounds stop
    at the adaptation boundary)."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    outs = []
    for depth in (1, 2, 3, 5, 0):
        mc = pkg.MCMC(model, g["data"], 1350.0, priors, q0, nsamples=47 if isinstance(priors, dict) else 23,
                      n_chains=24, verbose=False, seed=11, param_names=names, bounds=bounds, spec_depth=depth)
        assert mc.compat_adapt == isinstance(priors, dict)
        out = mc.sample(False)
        outs.append((out, mc.std2.copy(), mc.accepts.copy(), mc.stats, mc.checkpoint()["chol"]))
    ref = outs[0]
    assert ref[3]["nsolves_executed"] == ref[3]["nsolves"]
    if isinstance(priors, dict):                     # the proposal scale did move away from Vstart
        assert not np.any(ref[4][0] == mc.Vstart[0, 0])
    for depth, (out, s2, acc, stats, chol) in zip((2, 3, 5, 0), outs[1:]):
        assert np.array_equal(out, ref[0]), depth
        assert np.array_equal(s2, ref[1]), depth
        assert np.array_equal(acc, ref[2]), depth
        assert np.array_equal(chol, ref[4]), depth
        assert stats["nsolves"] == ref[3]["nsolves"]
        assert stats["nsolves_executed"] > stats["nsolves"]          # speculation did extra work
    assert 0 < ref[2].mean() < 1


def test_checkpoint_resume_continues_the_same_chains(cuda, pkg, tmp_path):
    """Counter-based RNG + saved (q, SSE, sigma^2, proposal, iteration): 30 iterations == 18 + resume + 12."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    q0 = np.linspace(900.0, 2000.0, 40)
    kw = dict(n_chains=40, verbose=False)
    full = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=30, seed=21, **kw)
    full.sample(False)
    chain_full = full.samples_device.cpu().numpy()              # [31, 1, 40]
    part1 = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=18, seed=21, **kw)
    part1.sample(False)
    fn = str(tmp_path / "ckpt.json")
    part1.checkpoint(fn)
    part2 = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=12, resume=fn, **kw)
    part2.sample(False)
    assert np.array_equal(part1.samples_device.cpu().numpy(), chain_full[:19])
    assert np.array_equal(part2.samples_device.cpu().numpy()[1:], chain_full[19:])
    assert np.array_equal(part2.std2_device.cpu().numpy()[1:], full.std2_device.cpu().numpy()[19:])


def test_rsf_driver_facade(cuda, pkg, capsys, tmp_path, monkeypatch):
    """main.py's flow through the RSF facade: batched data generation, JSON round trip, one MCMC per Dc,
    inference() returns elapsed seconds (q11)."""
    monkeypatch.chdir(tmp_path)
    problem = pkg.RSF(number_slip_values=3, lowest_slip_value=500.0, largest_slip_value=2500.0, qstart=1000.0,
                      qpriors=["Uniform", 0.0, 10000.0])
    assert np.array_equal(problem.dc_list, [500.0, 1500.0, 2500.0]) and problem.num_dc == 3
    problem.model = pkg.RateStateModel(number_time_steps=500)
    np.random.seed(1)
    problem.data = problem.generate_time_series()
    assert problem.data.shape == (1500,) and problem.model.Dc == 2500.0
    # same data as three separate evaluate() calls with the same global seed (reference layout)
    np.random.seed(1)
    m2 = pkg.RateStateModel(number_time_steps=500)
    ref = []
    for dc in problem.dc_list:
        m2.Dc = dc
        ref.append(m2.evaluate()[2])
    assert np.allclose(problem.data, np.concatenate(ref), rtol=1e-12, atol=0)
    problem.format = "json"
    problem.mcmc_kwargs = {"verbose": False, "seed": 3}
    elapsed = problem.inference(nsamples=40)
    assert isinstance(elapsed, float) and elapsed > 0
    out = capsys.readouterr().out
    assert "--- Dc is 500.0 ---" in out and "--- Dc is 2500.0 ---" in out
    assert (tmp_path / "data.json").exists()
    for dc in problem.dc_list:
        r = problem.results[dc]
        assert r["samples"].shape == (1, 21)
    # the 1500 chain should sit much closer to 1500 than the 500 chain
    assert abs(problem.results[1500.0]["samples"].mean() - 1500.0) < 300.0
    problem.format = "mysql"
    with pytest.raises(NotImplementedError):
        problem.prepare_data(problem.data)


def test_joint_abdc_posterior_matches_cpu_chain(cuda, pkg, orc):
    """d = 3 (a, b, Dc) with pooled adaptive covariance (extension: the reference is d = 1 only).  Oracle for
    the posterior: an independent CPU random-walk Metropolis (NumPy + the C forward model) with the same
    likelihood, sigma^2 Gibbs step and box prior.  Posterior means must agree within Monte-Carlo error."""
    rng = np.random.default_rng(42)
    truth = (0.011, 0.014, 1325.0)
    om = orc.make_model(Dc=truth[2], a=truth[0], b=truth[1])
    _, acc_true, _ = orc.forward(om)
    data = acc_true + np.abs(acc_true) * rng.standard_normal(acc_true.size)
    bounds = np.array([[0.0100, 0.0120], [0.0130, 0.0150], [800.0, 2200.0]])
    n, n0 = data.size, 0.01

    def sse(q):
        m = orc.make_model(Dc=q[2], a=q[0], b=q[1])
        return orc.sse(orc.forward(m)[1], data)

    # ---- CPU reference chain ----
    q = np.array(truth)
    ss = sse(q)
    s2 = ss / (n - 3)
    step = np.array([2.5e-4, 2.5e-4, 60.0])
    draws = []
    for it in range(6000):
        qn = q + step * rng.standard_normal(3)
        if np.all((qn > bounds[:, 0]) & (qn < bounds[:, 1])):
            ssn = sse(qn)
            if min(0.0, 0.5 * (ss - ssn) / s2) > np.log(rng.random()):
                q, ss = qn, ssn
        s2 = 1.0 / (rng.gamma(0.5 * (n0 + n)) / (0.5 * (n0 * s2 + ss)))
        if it >= 1000:
            draws.append(q.copy())
    draws = np.array(draws)
    cpu_mean, cpu_sd = draws.mean(axis=0), draws.std(axis=0)
    # batch-means MCSE of the CPU chain
    bm = draws[: (len(draws) // 50) * 50].reshape(50, -1, 3).mean(axis=1)
    cpu_mcse = bm.std(axis=0, ddof=1) / np.sqrt(50)

    # ---- GPU: 1024 chains, pooled adaptive Metropolis ----
    c = 1024
    q0 = np.stack([rng.uniform(0.0105, 0.0115, c), rng.uniform(0.0135, 0.0145, c), rng.uniform(1000.0, 1800.0, c)], axis=1)
    mc = pkg.MCMC(pkg.RateStateModel(), data, truth[2], ["Uniform", 0.0, 1e4], q0, nsamples=500, n_chains=c,
                  verbose=False, seed=8, param_names=("a", "b", "Dc"), bounds=bounds, adapt="pooled", adapt_start=60)
    out = mc.sample(False)                                     # [c, 3, 251]
    assert out.shape == (c, 3, 251) and len(mc.adapt_history) > 10
    chain_means = out.mean(axis=2)                             # [c, 3]
    gpu_mean = chain_means.mean(axis=0)
    gpu_mcse = chain_means.std(axis=0, ddof=1) / np.sqrt(c)
    gpu_sd = out.transpose(1, 0, 2).reshape(3, -1).std(axis=1)
    for j in range(3):
        tol = 5.0 * np.hypot(cpu_mcse[j], gpu_mcse[j]) + 0.02 * cpu_sd[j]
        assert abs(gpu_mean[j] - cpu_mean[j]) < tol, (j, gpu_mean[j], cpu_mean[j], tol)
        assert 0.8 < gpu_sd[j] / cpu_sd[j] < 1.25, (j, gpu_sd[j], cpu_sd[j])
    assert 0.1 < mc.acceptance_ratio.mean() < 0.7


@pytest.mark.timeout(120)
def test_non_finite_state_terminates(cuda, pkg):
    """A NaN in the observed series makes every sum of squares NaN: nothing can be accepted, and both the
    one-thread-per-chain and the speculative kernel must still finish (no spin on an undecidable round)."""
    data = np.zeros(500)
    data[17] = np.nan
    for depth in (1, 0):
        mc = pkg.MCMC(pkg.RateStateModel(), data, 1000.0, ["Uniform", 0.0, 1e4], 1000.0, nsamples=8, n_chains=8,
                      verbose=False, seed=1, spec_depth=depth)
        out = mc.sample(False)
        assert out.shape == (8, 1, 5)
        assert np.all(out == 1000.0) and mc.accepts.sum() == 0


def test_init_workspace_cache_and_trim(cuda, pkg):
    """rsfm_init's workspace is recycled between samplers (and between sizes / d); results do not
    depend on whether the buffer came from the cache, and rsfm_trim releases everything."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    lib = pkg._lib.load()

    def run(n_chains):
        return pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], 1000.0, n_chains=n_chains, nsamples=10,
                        verbose=False, seed=5).sample(False)

    assert lib.rsfm_trim() == 0
    first = run(96)                    # fresh allocation
    again = run(96)                    # same buffer from the cache
    small = run(32)                    # a larger cached buffer serves a smaller request
    big = run(160)                     # cache miss: a new, larger buffer replaces or joins it
    assert np.array_equal(first, again)
    assert np.array_equal(small, first[:32]) and np.array_equal(big[:96], first)
    assert lib.rsfm_trim() == 0
    assert np.array_equal(run(96), first)
    out = model.evaluate_batch(np.array([800.0, 1350.0]), data=g["data"])    # nominal-table cache rebuilt after trim
    assert np.isfinite(out["sse"].cpu().numpy()).all()


def test_shard_invariance_at_cfg5_shard_size(cuda, pkg):
    """131,072 chains (one GPU's share of cfg 5): the chains do not depend on how they are split over
    samplers -- two half-size samplers with the right global chain ids reproduce the full run bit for
    bit -- and every chain differs from its neighbour."""
    m = pkg.RateStateModel()
    m.Dc = 1325.0
    np.random.seed(8)
    _, _, data = m.evaluate()
    c = 131072
    q0 = np.random.default_rng(2).uniform(200.0, 5000.0, c)
    kw = dict(nsamples=6, verbose=False, seed=99)
    full = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0, n_chains=c, **kw)
    out = full.sample(False)
    assert out.shape == (c, 1, 4) and full.stats["failed_chains"] == 0
    h = c // 2
    lo = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0[:h], n_chains=h, chain_id0=0, **kw).sample(False)
    hi = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0[h:], n_chains=h, chain_id0=h, **kw).sample(False)
    assert np.array_equal(out[:h], lo) and np.array_equal(out[h:], hi)
    moved = np.abs(out[:, 0, -1] - q0) > 0
    assert 0.2 < moved.mean() <= 1.0          # most chains accepted something within the first iterations


def test_out_of_bounds_walk_with_idle_warps_and_pooled_sums(cuda, pkg):
    """d = 3 sequential kernel at a size whose last 64-thread block holds one chain and one empty warp:
    lanes that skip ahead through out-of-bounds iterations must leave every output row written, the
    pooled sums complete, and the chains independent of the launch partition (8 = 3 + 3 + 2 iterations)."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    c = 4737
    rng = np.random.default_rng(4)
    q0 = np.stack([rng.uniform(0.0104, 0.0106, c), rng.uniform(0.0144, 0.0146, c), rng.uniform(1180.0, 1220.0, c)], axis=1)
    kw = dict(n_chains=c, verbose=False, seed=21, param_names=("a", "b", "Dc"), spec_depth=1,
              bounds=[[0.0100, 0.0110], [0.0140, 0.0150], [1100.0, 1300.0]])
    one = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=8, **kw)
    out = one.sample(False)
    full = one.samples_device.cpu().numpy()                    # [9, 3, C] incl. the start values
    assert np.isfinite(full).all() and np.all(full[:, 0] > 0.0100) and np.all(full[:, 2] < 1300.0)
    acc = one.accept_device.cpu().numpy()
    moved = np.any(full[1:] != full[:-1], axis=1)              # [8, C]
    assert np.array_equal(moved, acc.astype(bool))
    assert one.stats["nsolves"] < 8 * c + 4 * c               # out-of-bounds proposals cost no solve
    # pooled adaptation path: launches of adapt_interval iterations, sums over every iteration of every chain
    pooled = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=8, adapt="pooled",
                      adapt_start=1000, adapt_interval=3, **kw)
    out_p = pooled.sample(False)
    assert np.array_equal(out_p, out)                          # no adaptation before adapt_start: same chains


@pytest.mark.parametrize("spec_depth", [0, 1])
def test_stiff_velocity_step_chains_replay_through_oracle(cuda, pkg, orc, spec_depth):
    """cfg-4 style at reduced size (velocity steps x10, Dc ~ 0.05: stability-limited DOP853): the stiff kernel
    variant inside the sampler -- speculative kernel and, with cfg.spec_depth = 1, the sequential one.  The draws the
    kernel used are replayed on the CPU oracle: same accept / reject decision at every step."""
    import ctypes as C
    torch = cuda
    n, t_end, period, factor = 400, 40.0, 10.0, 10.0
    om = orc.make_model(Dc=0.05, number_time_steps=n, end_time=t_end, loading=orc.LOAD_VSTEP, vstep_period=period,
                        vstep_factor=factor)
    _, acc_true, _ = orc.forward(om)
    rng = np.random.default_rng(11)
    data = acc_true + 0.2 * np.abs(acc_true) * rng.standard_normal(acc_true.size)
    lib = pkg._lib.load()
    model = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    model.loading, model.vstep_period, model.vstep_factor = "vstep", period, factor
    cfg = model.to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = 1, 3, spec_depth
    cfg.lo[0], cfg.hi[0] = 0.03, 0.09
    c, ns = 32, 12
    dev = torch.device("cuda", 0)
    q0 = torch.full((1, c), 0.06, dtype=torch.float64, device=dev)
    data_t = torch.from_numpy(data).to(dev)
    h = lib.rsfm_create(C.byref(cfg), c, 77, 0)
    assert h
    try:
        pkg._lib.check(lib.rsfm_init(h, q0.data_ptr(), data_t.data_ptr(), None))
        samples = torch.empty((ns, 1, c), dtype=torch.float64, device=dev)
        s2 = torch.empty((ns, c), dtype=torch.float64, device=dev)
        acc = torch.empty((ns, c), dtype=torch.uint8, device=dev)
        draws = torch.empty((ns, 3, c), dtype=torch.float64, device=dev)
        pkg._lib.check(lib.rsfm_run(h, ns, samples.data_ptr(), s2.data_ptr(), acc.data_ptr(), draws.data_ptr(), None))
        torch.cuda.synchronize()
    finally:
        lib.rsfm_destroy(h)
    samples, s2, acc, draws = (x.cpu().numpy() for x in (samples, s2, acc, draws))
    assert 0 < acc.mean() < 1
    for ch in (0, 5, 31):
        prop, u, gam = draws[:, 0, ch], draws[:, 1, ch], draws[:, 2, ch]
        chain_o, s2_o, acc_o, _, _ = orc.chain_replay(om, data, 0.06, 0.03, 0.09, 3, ns, prop, np.nan_to_num(u, nan=0.5), gam)
        assert np.array_equal(acc[:, ch], acc_o)
        assert np.array_equal(samples[:, 0, ch], chain_o[1:])
        assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-4, atol=0)          # SSE to ~1e-6 in the stiff regime


def _run_with_draws(torch, pkg, cfg, c, seed, id0, q0, data, ns_list):
    """rsfm_create / init / run through the C ABI with the draws dumped; returns per-iteration arrays."""
    import ctypes as C
    lib = pkg._lib.load()
    d = cfg.n_params
    q0_t = torch.from_numpy(np.broadcast_to(np.asarray(q0, dtype=np.float64).reshape(d, -1), (d, c)).copy()).cuda()
    data_t = torch.from_numpy(np.ascontiguousarray(data)).cuda()
    h = lib.rsfm_create(C.byref(cfg), c, seed, id0)
    assert h, lib.rsfm_last_error()
    outs = []
    try:
        pkg._lib.check(lib.rsfm_init(h, q0_t.data_ptr(), data_t.data_ptr(), None))
        depth = lib.rsfm_spec_depth(h)
        for ns in ns_list:
            samples = torch.empty((ns, d, c), dtype=torch.float64, device="cuda")
            s2 = torch.empty((ns, c), dtype=torch.float64, device="cuda")
            acc = torch.empty((ns, c), dtype=torch.uint8, device="cuda")
            draws = torch.empty((ns, d + 2, c), dtype=torch.float64, device="cuda")
            pkg._lib.check(lib.rsfm_run(h, ns, samples.data_ptr(), s2.data_ptr(), acc.data_ptr(), draws.data_ptr(), None))
            torch.cuda.synchronize()
            outs.append([x.cpu().numpy() for x in (samples, s2, acc, draws)])
        tot = (C.c_uint64 * 9)()
        pkg._lib.check(lib.rsfm_get_totals(h, tot, None))
    finally:
        lib.rsfm_destroy(h)
    cat = [np.concatenate([o[i] for o in outs]) for i in range(4)]
    return cat + [depth, list(tot)]


@pytest.mark.parametrize("spec_depth", [1, 3])
def test_joint_abdc_chains_replay_step_for_step_through_oracle(cuda, pkg, orc, spec_depth):
    """d = 3 (a, b, Dc), Philox-driven: the SEQUENTIAL kernel -- whose lanes walk ahead through their own
    out-of-bounds iterations between solves (rsfm_kernels.cu, SKIP_T) -- and the speculative one, replayed on
    the CPU oracle with the draws they report: the same accept / reject decision at every step of every checked
    chain, the same samples, sigma^2 to 1e-8; and the proposals are q + L z with the documented Philox stream
    and the Cholesky factor of Vstart."""
    torch = cuda
    rng = np.random.default_rng(12)
    om = orc.make_model(Dc=1325.0)
    truth = orc.forward(om)[1]
    data = truth + np.abs(truth) * rng.standard_normal(truth.size)
    cfg = pkg.RateStateModel().to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = 3, 3, spec_depth
    c, seed, id0, ns = 70, 2025, 4000, 36                # 70 chains: 2 full warps + a ragged one
    q0 = np.array([0.011, 0.014, 1300.0])
    # rsfm_init caps every marginal proposal s.d. at 1/20 of the prior width (no reference behaviour exists for
    # d = 3), so a box centred on the start value is hardly ever left; the start value sits half an s.d. from the
    # lower corner instead: a third to a half of the proposals leave the box
    width = np.array([0.0012, 0.0012, 400.0])
    lo = q0 - width / 40.0
    hi = lo + width
    for j in range(3):
        cfg.lo[j], cfg.hi[j] = lo[j], hi[j]
    samples, s2, acc, draws, depth, tot = _run_with_draws(torch, pkg, cfg, c, seed, id0, q0, data, [20, 16])
    assert depth == (0 if spec_depth == 1 else 3)
    oob = np.isnan(draws[:, 3])
    assert 0.15 < oob.mean() < 0.9 and 0 < acc.mean() < 1
    in_box = np.all((draws[:, :3] > lo[None, :, None]) & (draws[:, :3] < hi[None, :, None]), axis=1)
    assert np.array_equal(in_box, ~oob)
    assert tot[0] == in_box.sum() + 4 * c                # one solve per in-bounds proposal (+ the 1 + d set-up solves)
    for ch in (0, 1, 31, 32, 63, 64, 69):
        prop = draws[:, :3, ch]
        chain_o, s2_o, acc_o, nsolves = orc.chain_replay_nd(om, data, q0, lo, hi, 3, ns, prop,
                                                            np.nan_to_num(draws[:, 3, ch], nan=0.5), draws[:, 4, ch])
        assert np.array_equal(acc[:, ch], acc_o), ch
        assert np.array_equal(samples[:, :, ch], chain_o[1:]), ch
        assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-8, atol=0)
    # the proposals themselves: q_cur + L z, z = the chain's Philox normals, L = chol(Vstart) (no adaptation)
    lib = pkg._lib.load()
    hook = torch.empty((ns, 6, c), dtype=torch.float64, device="cuda")
    pkg._lib.check(lib.rsfm_philox_draws(seed, id0, c, 0, ns, 0.5 * (0.01 + 500), hook.data_ptr(), None))
    hook = hook.cpu().numpy()
    mc = pkg.MCMC(pkg.RateStateModel(), data, 1325.0, ["Uniform", 0.0, 1e4], q0, nsamples=2, n_chains=c, verbose=False,
                  seed=seed, chain_id0=id0, param_names=("a", "b", "Dc"), bounds=np.stack([lo, hi], axis=1), spec_depth=1)
    mc.sample(False)
    L = np.zeros((3, 3))
    L[np.tril_indices(3)] = mc.checkpoint()["chol"][:, 0]                       # row-major lower triangle
    assert np.allclose(L @ L.T, mc.Vstart, rtol=1e-12)
    cur = np.concatenate([np.broadcast_to(q0.reshape(1, 3, 1), (1, 3, c)), samples[:-1]])
    z = hook[:, :3]                                                             # [ns, 3, c]
    expect = cur + np.einsum("ij,njc->nic", L, z)
    assert np.allclose(draws[:, :3], expect, rtol=1e-9, atol=0)
    assert np.array_equal(draws[:, 3][~oob], hook[:, 3][~oob]) and np.array_equal(draws[:, 4], hook[:, 4])


@pytest.mark.parametrize("d,c,spec_depth", [(1, 37, 1), (1, 1, 0), (1, 200, 0), (3, 45, 1), (1, 4100, 1)])
def test_series_of_513_to_1024_points(cuda, pkg, orc, d, c, spec_depth):
    """number_time_steps = 1000 (the value of the reference's docstring examples): a series of two resident
    tiles.  Warps whose lanes are all rejected early, out of bounds or inactive leave the output loop before the
    second tile -- sequential d = 1, the d = 3 out-of-bounds walk, the speculative kernel with ONE chain (idle
    tree lanes), and chain counts that are no multiple of the block size -- and the chains must still be the
    oracle's, step for step."""
    torch = cuda
    n, t_end = 1000, 100.0
    rng = np.random.default_rng(3)
    om = orc.make_model(Dc=1325.0, number_time_steps=n, end_time=t_end)
    truth = orc.forward(om)[1]
    data = truth + np.abs(truth) * rng.standard_normal(n)
    model = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    cfg = model.to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = d, 3, spec_depth
    if d == 1:
        lo, hi, q0 = np.array([1250.0]), np.array([1420.0]), np.array([1330.0])
    else:
        lo, hi, q0 = np.array([0.0104, 0.0134, 1100.0]), np.array([0.0116, 0.0146, 1500.0]), np.array([0.011, 0.014, 1300.0])
    for j in range(d):
        cfg.lo[j], cfg.hi[j] = lo[j], hi[j]
    ns = 14
    samples, s2, acc, draws, depth, tot = _run_with_draws(torch, pkg, cfg, c, 5, 0, q0, data, [ns])
    assert (depth >= 2) == (spec_depth == 0)
    assert tot[5] > 0 or c == 1                           # solves were stopped early (rejection certain)
    for ch in sorted({0, c // 2, c - 1}):
        chain_o, s2_o, acc_o, _ = orc.chain_replay_nd(om, data, q0, lo, hi, 3, ns, draws[:, :d, ch],
                                                      np.nan_to_num(draws[:, d, ch], nan=0.5), draws[:, d + 1, ch])
        assert np.array_equal(acc[:, ch], acc_o), ch
        assert np.array_equal(samples[:, :, ch], chain_o[1:]), ch
        assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-8, atol=0)


def test_mu_observable_end_to_end(cuda, pkg, orc):
    """RSFM_OBS_MU (north_star's "observed friction series"): the solver scores mu_k, the series
    the reference integrates and stores (RateStateModel.py:367, 385): trajectories against the oracle's mu,
    SSE against the oracle's, and Philox-driven chains replayed on the oracle with observable = mu."""
    torch = cuda
    rng = np.random.default_rng(8)
    om = orc.make_model(Dc=1325.0, observable=orc.OBS_MU)
    _, mu_true, _, _, _, _ = orc.forward(orc.make_model(Dc=1325.0), full=True)
    sig = 2e-7
    data = mu_true + sig * rng.standard_normal(mu_true.size)
    m = pkg.RateStateModel()
    m.observable = "mu"
    dcs = np.array([1.0, 60.0, 300.0, 1000.0, 1325.0, 5000.0])
    out = m.evaluate_batch(dcs, data=data)
    mu_g = out["acc"].t().cpu().numpy()
    for i, dc in enumerate(dcs):
        _, mu_o, _, _, _, _ = orc.forward(orc.make_model(Dc=dc), full=True)
        assert mu_g[i, 0] == 0.6
        tol = (1e-6 if dc < 50 else 1e-9) * np.max(np.abs(mu_o - 0.6)) + 2e-16
        assert np.max(np.abs(mu_g[i] - mu_o)) <= tol, dc
    sse_o, _, _ = orc.forward_batch(om, dcs, data=data)
    assert np.allclose(out["sse"].cpu().numpy()[1:], sse_o[1:], rtol=1e-8, atol=0)
    # evaluate() returns the observable in slot 1, like the reference's protocol (MCMC.py:127)
    m.Dc = 1325.0
    t, mu1, mu_noise = m.evaluate()
    assert np.array_equal(mu1, mu_g[4]) and mu_noise.shape == mu1.shape and mu_noise[0] == 0.6
    # chains on the friction series, replayed through the oracle
    cfg = m.to_cfg()
    cfg.n_params, cfg.n_prior_len = 1, 3
    cfg.lo[0], cfg.hi[0] = 1200.0, 1450.0
    for spec_depth in (1, 0):
        cfg.spec_depth = spec_depth
        samples, s2, acc, draws, depth, tot = _run_with_draws(torch, pkg, cfg, 48, 31, 100, np.array([1300.0]), data, [30])
        assert 0 < acc.mean() < 1
        for ch in (0, 20, 47):
            chain_o, s2_o, acc_o, _ = orc.chain_replay_nd(om, data, [1300.0], [1200.0], [1450.0], 3, 30, draws[:, :1, ch],
                                                          np.nan_to_num(draws[:, 1, ch], nan=0.5), draws[:, 2, ch])
            assert np.array_equal(acc[:, ch], acc_o), (spec_depth, ch)
            assert np.array_equal(samples[:, 0, ch], chain_o[1:, 0])
            assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-7, atol=0)
    # the public sampler on the friction series recovers Dc
    mc = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], 1000.0, nsamples=400, n_chains=256, verbose=False, seed=2)
    post = mc.sample(False)
    assert abs(post[:, 0, :].mean() - 1325.0) < 4 * post[:, 0, :].std() and post[:, 0, :].std() < 100.0


def test_device_pooled_update_matches_host_algebra(cuda, pkg):
    """rsfm_pooled_partials / rsfm_pooled_update (sharding-invariant partial sums, moments, closed-form Cholesky,
    install -- all on the device) against the NumPy restatement in adaptation.py, d = 1 and d = 3; and the
    partial rows of a sampler are the same bits whether it holds the whole range of chains or a shard of it."""
    import ctypes as C
    import importlib
    torch = cuda
    ad = importlib.import_module("bayesian-markov-chain-monte-carlo_b200.adaptation")
    lib = pkg._lib.load()
    g = load_golden("sse_grid.json")
    data_t = torch.from_numpy(g["data"]).cuda()
    rows = pkg._lib.POOL_ROWS
    for d, q0, lo, hi in ((1, [1300.0], [0.0], [1e4]),
                          (3, [0.011, 0.014, 1300.0], [0.0100, 0.0130, 800.0], [0.0120, 0.0150, 2200.0])):
        cfg = pkg.RateStateModel().to_cfg()
        cfg.n_params, cfg.n_prior_len, cfg.adapt_mode = d, 3, pkg._lib.ADAPT_POOLED
        for j in range(d):
            cfg.lo[j], cfg.hi[j] = lo[j], hi[j]
        tri = d * (d + 1) // 2

        def run(c, id0, ns):
            q = torch.from_numpy(np.repeat(np.array(q0).reshape(d, 1), c, axis=1).copy()).cuda()
            h = lib.rsfm_create(C.byref(cfg), c, 77, id0)
            assert h
            try:
                pkg._lib.check(lib.rsfm_init(h, q.data_ptr(), data_t.data_ptr(), None))
                samples = torch.empty((ns, d, c), dtype=torch.float64, device="cuda")
                pkg._lib.check(lib.rsfm_run(h, ns, samples.data_ptr(), None, None, None, None))
                ng = lib.rsfm_pooled_groups(h)
                part = torch.full((ng, rows), -1.0, dtype=torch.float64, device="cuda")
                pkg._lib.check(lib.rsfm_pooled_partials(h, part.data_ptr(), 1, None))
                mom = torch.zeros(1 + d + tri, dtype=torch.float64, device="cuda")
                fac = torch.zeros(1 + tri, dtype=torch.float64, device="cuda")
                pkg._lib.check(lib.rsfm_pooled_update(h, part.data_ptr(), ng, mom.data_ptr(), 1, 1, fac.data_ptr(), None))
                chol = torch.empty((tri, c), dtype=torch.float64, device="cuda")
                pkg._lib.check(lib.rsfm_get_state(h, None, None, None, chol.data_ptr(), None, None, None, None, None))
                torch.cuda.synchronize()
                return samples.cpu().numpy(), part.cpu().numpy(), mom.cpu().numpy(), fac.cpu().numpy(), chol.cpu().numpy()
            finally:
                lib.rsfm_destroy(h)

        c, ns = 3072, 12
        samples, part, mom, fac, chol = run(c, 2048, ns)
        assert part.shape == (3, rows) and np.all(part[:, 1 + d + tri:] == 0.0)
        x = samples.transpose(0, 2, 1).reshape(-1, d)                      # all draws of all chains
        assert mom[0] == ns * c
        assert np.allclose(mom[1:1 + d], x.sum(axis=0), rtol=1e-12)
        sec = x.T @ x
        assert np.allclose(mom[1 + d:], sec[np.tril_indices(d)], rtol=1e-12)
        want = ad.proposal_from_suffstats(mom, d)
        assert fac[0] == 1.0
        if d == 1:
            assert np.allclose(fac[1:], want, rtol=1e-7)
        else:
            # compared as covariances: after 12 iterations from one start value the three parameters are almost
            # perfectly correlated, and the last pivot l22^2 = v22 - l20^2 - l21^2 is 2e-10 of v22 -- the factor's
            # last entries carry the rounding of the raw moments amplified by that ratio (7e-4 observed), in the
            # host algebra just as on the device; L L^T is what the proposals see and is well conditioned
            def cov_of(f):
                low = np.zeros((d, d))
                low[np.tril_indices(d)] = f
                return low @ low.T
            v_dev, v_host = cov_of(fac[1:]), cov_of(want)
            scale = np.sqrt(np.outer(np.diag(v_host), np.diag(v_host)))
            assert np.max(np.abs(v_dev - v_host) / scale) < 1e-9
            assert np.allclose(fac[1:4], want[:3], rtol=1e-7)
        assert np.all(chol == fac[1:, None])                               # installed for every chain
        # shard invariance of the partial rows: chains [2048, 5120) as [2048, 4096) + [4096, 5120)
        _, part_a, _, _, _ = run(2048, 2048, ns)
        _, part_b, _, _, _ = run(1024, 4096, ns)
        assert np.array_equal(np.concatenate([part_a, part_b]), part)
        # a range that does not start on a group boundary: more groups, same total
        _, part_u, mom_u, _, _ = run(3072, 2048 + 100, ns)
        assert part_u.shape[0] == 4 and part_u[:, 0].sum() == ns * 3072
    # degenerate moments: identical samples -> no factor, the proposal stays
    cfg = pkg.RateStateModel().to_cfg()
    cfg.n_params, cfg.adapt_mode = 1, pkg._lib.ADAPT_POOLED
    q = torch.full((1, 64), 1300.0, dtype=torch.float64, device="cuda")
    h = lib.rsfm_create(C.byref(cfg), 64, 1, 0)
    try:
        pkg._lib.check(lib.rsfm_init(h, q.data_ptr(), data_t.data_ptr(), None))
        before = torch.empty((1, 64), dtype=torch.float64, device="cuda")
        pkg._lib.check(lib.rsfm_get_state(h, None, None, None, before.data_ptr(), None, None, None, None, None))
        mom = torch.tensor([640.0, 640.0 * 1300.0, 640.0 * 1300.0 ** 2], dtype=torch.float64, device="cuda")
        fac = torch.ones(2, dtype=torch.float64, device="cuda")
        pkg._lib.check(lib.rsfm_pooled_update(h, None, 0, mom.data_ptr(), 0, 1, fac.data_ptr(), None))
        after = torch.empty((1, 64), dtype=torch.float64, device="cuda")
        pkg._lib.check(lib.rsfm_get_state(h, None, None, None, after.data_ptr(), None, None, None, None, None))
        torch.cuda.synchronize()
        assert fac[0].item() == 0.0 and torch.equal(before, after)
    finally:
        lib.rsfm_destroy(h)


def test_pooled_sampler_single_chain_and_checkpoint(cuda, pkg, tmp_path):
    """adapt='pooled' with n_chains = 1 (the draws buffer is filled: model.Dc and the per-iteration printout are
    the chain's own values), and exact continuation from a checkpoint on an adaptation boundary: the pooled
    moments and the gathered rows of the last interval are part of the state."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    one = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], 1000.0, nsamples=40, verbose=False, seed=3,
                   adapt="pooled", adapt_start=20)
    out = one.sample(False)
    assert out.shape == (1, 21) and np.all(np.isfinite(out)) and 0.0 < float(model.Dc[0]) < 1e4
    kw = dict(n_chains=2048, verbose=False, adapt="pooled", adapt_start=20, adapt_interval=10)
    q0 = np.linspace(900.0, 2000.0, 2048)
    full = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=70, seed=9, **kw)
    full.sample(False)
    assert full.stats["n_adaptations"] >= 4 and len(full.adapt_history) == full.stats["n_adaptations"]
    part1 = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=40, seed=9, **kw)
    part1.sample(False)
    fn = str(tmp_path / "pooled.json")
    part1.checkpoint(fn)
    part2 = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=30, resume=fn, **kw)
    part2.sample(False)
    a = full.samples_device.cpu().numpy()
    assert np.array_equal(part1.samples_device.cpu().numpy(), a[:41])
    assert np.array_equal(part2.samples_device.cpu().numpy()[1:], a[41:])
    # a checkpoint in the middle of an adaptation interval is refused, not silently inexact
    mid = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=25, seed=9, **kw)
    mid.sample(False)
    with pytest.raises(ValueError, match="adaptation boundary"):
        pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=10, resume=mid.checkpoint(), **kw).sample(False)


def test_compat_checkpoint_mid_window_resumes_exactly(cuda, pkg):
    """dict priors (the reference's windowed adaptation): the 10-sample ring is part of the checkpoint, so a
    resume in the middle of a window continues the same chains, bit for bit."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    pri = {1: 0.0, 2: 1e4}
    kw = dict(n_chains=24, verbose=False)
    full = pkg.MCMC(model, g["data"], 1350.0, pri, 1000.0, nsamples=47, seed=4, **kw)
    full.sample(False)
    p1 = pkg.MCMC(model, g["data"], 1350.0, pri, 1000.0, nsamples=23, seed=4, **kw)
    p1.sample(False)
    p2 = pkg.MCMC(model, g["data"], 1350.0, pri, 1000.0, nsamples=24, resume=p1.checkpoint(), **kw)
    p2.sample(False)
    a = full.samples_device.cpu().numpy()
    assert np.array_equal(p1.samples_device.cpu().numpy(), a[:24])
    assert np.array_equal(p2.samples_device.cpu().numpy()[1:], a[24:])
    assert np.array_equal(p2.checkpoint()["chol"], full.checkpoint()["chol"])


@pytest.mark.parametrize("spec_depth", [1, 0])
def test_slip_law_and_tabulated_load_chains_replay_through_oracle(cuda, pkg, orc, spec_depth):
    """Sampler level for the SURVEY 8f.4 extensions: Philox-driven chains under Ruina's slip law with a tabulated
    load (sequential and speculative kernel) replayed on the CPU oracle with the draws they report -- the same
    accept / reject decision at every step, the same samples, sigma^2 to 1e-8."""
    torch = cuda
    tt = np.arange(501) * 0.1
    tab = 0.5 * np.sin(0.7 * tt) * np.exp(-tt / 30.0) + 0.3 * (tt > 20.0)
    om = orc.make_model(Dc=1325.0, state_law=orc.LAW_SLIP, loading=orc.LOAD_TABLE, load_table=tab, load_dt=0.1)
    rng = np.random.default_rng(21)
    truth = orc.forward(om)[1]
    data = truth + np.abs(truth) * rng.standard_normal(truth.size)
    model = pkg.RateStateModel()
    model.state_law, model.loading, model.load_table, model.load_dt = "slip", "table", tab, 0.1
    cfg = model.to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = 1, 3, spec_depth
    cfg.lo[0], cfg.hi[0] = 0.0, 1e4
    c, ns = 40, 24
    samples, s2, acc, draws, depth, tot = _run_with_draws(torch, pkg, cfg, c, 11, 500, [1100.0], data, [ns])
    assert (depth >= 2) == (spec_depth == 0)
    assert 0 < acc.mean() < 1
    for ch in (0, 7, 33, 39):
        chain_o, s2_o, acc_o, _, _ = orc.chain_replay(om, data, 1100.0, 0.0, 1e4, 3, ns, draws[:, 0, ch],
                                                      np.nan_to_num(draws[:, 1, ch], nan=0.5), draws[:, 2, ch])
        assert np.array_equal(acc[:, ch], acc_o), ch
        assert np.array_equal(samples[:, 0, ch], chain_o[1:]), ch
        assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-8, atol=0)


@pytest.mark.parametrize("d", [1, 3])
def test_chain_groups_give_the_same_chains(cuda, pkg, d):
    """Pooled adaptation with the chains served by four launches on the sampler's own streams (rsfm_chain_groups,
    include/rsfm.h) against one launch: the same samples, sigma^2, accept flags and pooled factors, bit for bit --
    also when the shard does not start at chain 0 and through a checkpoint taken while the groups are in flight."""
    rng = np.random.default_rng(8)
    g = load_golden("sse_grid.json")
    c, ns = 32768, 50
    kw = dict(nsamples=ns, n_chains=c, seed=9, chain_id0=3 * 1024, verbose=False, adapt="pooled", adapt_start=20)
    if d == 1:
        q0 = rng.uniform(900.0, 2000.0, c)
    else:
        q0 = np.stack([rng.uniform(0.0105, 0.0115, c), rng.uniform(0.0135, 0.0145, c), rng.uniform(1000.0, 1800.0, c)], axis=1)
        kw.update(param_names=("a", "b", "Dc"), bounds=[[0.0100, 0.0120], [0.0130, 0.0150], [800.0, 2200.0]])
    res = {}
    for groups in (1, 0):
        m = pkg.RateStateModel()
        m.chain_groups = groups
        mc = pkg.MCMC(m, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, **kw)
        out = mc.sample(False)
        res[groups] = (np.array(out), np.array(mc.std2), np.array(mc.accepts),
                       [(int(e), [float(x) for x in f]) for e, f in mc.adapt_history], dict(mc.stats), mc.checkpoint())
    assert res[0][4].get("chain_groups") == 4 and res[1][4].get("chain_groups") == 1
    assert len(res[1][3]) >= 2 and res[0][3] == res[1][3]
    for i in range(3):
        assert np.array_equal(res[0][i], res[1][i]), i
    for k in ("nsolves", "nrhs", "nstep", "nsolves_stopped_early"):
        assert res[0][4][k] == res[1][4][k], k
    for k in ("q", "sse", "sigma2", "chol"):
        assert np.array_equal(np.asarray(res[0][5][k]), np.asarray(res[1][5][k])), k


def test_rsfm_join_orders_the_callers_stream_behind_the_chain_groups(cuda, pkg):
    """C-ABI contract of the chain groups (include/rsfm.h): with more than one group rsfm_run works on the sampler's
    own streams; rsfm_join (and every getter) orders the caller's stream behind it.  Outputs read behind the join
    equal those of the single-launch sampler, on the default stream and on a side stream."""
    import ctypes as C
    torch = cuda
    lib = pkg._lib.load()
    g = load_golden("sse_grid.json")
    data_t = torch.from_numpy(np.ascontiguousarray(g["data"])).cuda()
    c, ns = 16384, 10
    q0 = torch.from_numpy(np.random.default_rng(4).uniform(900.0, 2000.0, (1, c))).cuda()
    outs = {}
    for groups, use_side in ((1, False), (2, False), (2, True)):
        cfg = pkg.RateStateModel().to_cfg()
        cfg.adapt_mode, cfg.chain_groups = pkg._lib.ADAPT_POOLED, groups
        cfg.spec_depth = 1            # one thread per chain (16,384 chains would get two lanes each, and then no groups)
        st = torch.cuda.Stream() if use_side else torch.cuda.current_stream()
        with torch.cuda.stream(st):
            h = lib.rsfm_create(C.byref(cfg), c, 3, 0)
            assert h and lib.rsfm_chain_groups(h) == groups
            try:
                sp = st.cuda_stream
                pkg._lib.check(lib.rsfm_init(h, q0.data_ptr(), data_t.data_ptr(), sp))
                samples = torch.zeros((2 * ns, 1, c), dtype=torch.float64, device="cuda")
                acc = torch.zeros((2 * ns, c), dtype=torch.uint8, device="cuda")
                for part in range(2):
                    pkg._lib.check(lib.rsfm_run(h, ns, samples[part * ns:].data_ptr(), None, acc[part * ns:].data_ptr(), None, sp))
                pkg._lib.check(lib.rsfm_join(h, sp))
                host = samples.to("cpu", non_blocking=False)             # stream-ordered copy behind the join
                tot = (C.c_uint64 * 9)()
                pkg._lib.check(lib.rsfm_get_totals(h, tot, sp))
                outs[(groups, use_side)] = (host.numpy().copy(), acc.cpu().numpy().copy(), list(tot))
            finally:
                lib.rsfm_destroy(h)
    ref = outs[(1, False)]
    assert 0 < ref[1].mean() < 1
    for key in ((2, False), (2, True)):
        assert np.array_equal(outs[key][0], ref[0]) and np.array_equal(outs[key][1], ref[1]) and outs[key][2] == ref[2], key


@pytest.mark.parametrize("n, ns, c", [(2, 1, 1), (3, 2, 3), (5, 4, 33), (16, 6, 1)])
def test_smallest_grids_and_runs(cuda, pkg, orc, n, ns, c):
    """Edge sizes: a series of 2 .. 16 output points (one output interval at the least), one or two iterations
    (nburn = int(nsamples / 2) = 0), one chain or a ragged warp -- forward solve and chains against the oracle."""
    torch = cuda
    t_end = 0.1 * n
    om = orc.make_model(Dc=1325.0, number_time_steps=n, end_time=t_end)
    rng = np.random.default_rng(n)
    truth = orc.forward(om)[1]
    assert truth.size == n
    data = truth + (np.abs(truth) + 1e-4) * rng.standard_normal(n)
    m = pkg.RateStateModel(number_time_steps=n, end_time=t_end)
    out = m.evaluate_batch(np.array([300.0, 1325.0, 9000.0]), data=data)
    for i, dc in enumerate((300.0, 1325.0, 9000.0)):
        acc_o = orc.forward(orc.make_model(Dc=dc, number_time_steps=n, end_time=t_end))[1]
        assert np.max(np.abs(out["acc"][:, i].cpu().numpy() - acc_o)) <= 1e-9 * max(np.max(np.abs(acc_o)), 1e-30) + 1e-14
        assert out["sse"][i].item() == pytest.approx(float(np.sum((acc_o - data) ** 2)), rel=1e-9)
    cfg = m.to_cfg()
    cfg.n_params, cfg.n_prior_len = 1, 3
    cfg.lo[0], cfg.hi[0] = 0.0, 1e4
    if n <= 3:
        return                     # N - len(qpriors) <= 0: sigma^2_0 is not positive, in the reference as here
    samples, s2, acc, draws, depth, tot = _run_with_draws(torch, pkg, cfg, c, 5, 0, [1100.0], data, [ns])
    for ch in sorted({0, c - 1}):
        chain_o, s2_o, acc_o, _, _ = orc.chain_replay(om, data, 1100.0, 0.0, 1e4, 3, ns, draws[:, 0, ch],
                                                      np.nan_to_num(draws[:, 1, ch], nan=0.5), draws[:, 2, ch])
        assert np.array_equal(acc[:, ch], acc_o) and np.array_equal(samples[:, 0, ch], chain_o[1:])
        assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-8, atol=0)
    mc = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], 1100.0, nsamples=ns, verbose=False, seed=5)
    res = mc.sample(False)
    assert res.shape == (1, ns + 1 - int(ns / 2)) and mc.std2.shape == (ns + 1 - int(ns / 2),)


def test_round_packing_gives_the_same_chains(cuda, pkg, orc):
    """d = 3, a batch large enough for 128-thread blocks: with the in-bounds proposals of a block packed into its
    lowest threads at every solve round (rsf_mcmc_kernel<3, .., PACK>) and without -- the same samples, sigma^2,
    accept flags and work totals, bit for bit, on a ragged chain count with half of the proposals out of bounds; and
    a few chains of the packed run replayed on the oracle step for step."""
    torch = cuda
    rng = np.random.default_rng(31)
    om = orc.make_model(Dc=1325.0)
    truth = orc.forward(om)[1]
    data = truth + np.abs(truth) * rng.standard_normal(truth.size)
    c, ns = 19000 + 37, 20
    q0 = np.array([0.011, 0.014, 1300.0])
    width = np.array([0.0012, 0.0012, 400.0])
    lo = q0 - width / 40.0
    hi = lo + width
    res = {}
    for packing in (1, 0):
        m = pkg.RateStateModel()
        m.round_packing = packing
        cfg = m.to_cfg()
        cfg.n_params, cfg.n_prior_len, cfg.spec_depth = 3, 3, 1
        for j in range(3):
            cfg.lo[j], cfg.hi[j] = lo[j], hi[j]
        res[packing] = _run_with_draws(torch, pkg, cfg, c, 77, 2048, q0, data, [12, 8])
    for i in range(4):
        assert np.array_equal(res[0][i], res[1][i], equal_nan=True), i
    assert res[0][5] == res[1][5]                                   # work totals (solves, RHS, steps, early stops ...)
    samples, s2, acc, draws = res[0][:4]
    oob = np.isnan(draws[:, 3])
    assert 0.15 < oob.mean() < 0.9 and 0 < acc.mean() < 1
    for ch in (0, 127, 128, 9000, c - 1):
        chain_o, s2_o, acc_o, _ = orc.chain_replay_nd(om, data, q0, lo, hi, 3, ns, draws[:, :3, ch],
                                                      np.nan_to_num(draws[:, 3, ch], nan=0.5), draws[:, 4, ch])
        assert np.array_equal(acc[:, ch], acc_o), ch
        assert np.array_equal(samples[:, :, ch], chain_o[1:]), ch
        assert np.allclose(s2[:, ch], s2_o[1:], rtol=1e-8, atol=0)


@pytest.mark.parametrize("n_steps, t_end, c, loading", [(40, 4.0, 9600, "sine_decay"), (1100, 110.0, 6, "sine_decay"),
                                                       (2100, 210.0, 9500, "vstep")])
def test_two_lane_groups_and_streamed_series_in_the_speculative_kernel(cuda, pkg, n_steps, t_end, c, loading):
    """Round 2 widened the speculative kernel: two lanes per chain (9,473 .. 18,944 chains, where the one-thread-per-
    chain kernel runs one latency-bound warp per sub-partition) and a streamed series (n_out > 1,024; one-warp blocks,
    the tile barriers are warp-wide).  Chains must be those of the sequential kernel bit for bit."""
    model = pkg.RateStateModel(number_time_steps=n_steps, end_time=t_end)
    if loading == "vstep":
        model.loading, model.vstep_period, model.vstep_factor = "vstep", 50.0, 10.0
        truth, q0, lo, hi = 0.5, np.random.default_rng(2).uniform(0.3, 0.9, c), 0.05, 5.0
    else:
        truth, q0, lo, hi = 1325.0, np.random.default_rng(2).uniform(900.0, 2000.0, c), 0.0, 1e4
    model.Dc = truth
    np.random.seed(7)
    _, _, data = model.evaluate()
    outs = []
    for depth in (1, 0):
        mc = pkg.MCMC(model, data, truth, ["Uniform", lo, hi], q0, nsamples=8, n_chains=c, verbose=False, seed=5,
                      spec_depth=depth)
        out = mc.sample(False)
        outs.append((out, mc.std2.copy(), mc.accepts.copy(), mc.stats))
    assert outs[0][3]["nsolves_executed"] == outs[0][3]["nsolves"]
    assert outs[1][3]["nsolves_executed"] > outs[1][3]["nsolves"]            # the speculative kernel did run
    if loading == "vstep":
        # Stiff variant: whether a lane's step is scored by the fast or by the general-range stages depends on its
        # warp-mates (rsf_interval_general tries the fast step when ANY lane starts in range), the two agree to
        # rounding, and the stability-limited solve amplifies that to ~1e-6 .. 1e-4 of the sum of squares.
        # So the two kernels agree chain by chain at the level of decisions, not of bits.
        same = (outs[0][2] == outs[1][2]).all(axis=1)
        assert same.mean() > 0.98
        assert np.array_equal(outs[0][0][same], outs[1][0][same])
        assert np.allclose(outs[0][1][same], outs[1][1][same], rtol=1e-3, atol=0)
    else:
        assert outs[1][3]["nsolves"] == outs[0][3]["nsolves"]
        for a, b in zip(outs[0][:3], outs[1][:3]):
            assert np.array_equal(a, b)
    assert 0 < outs[0][2].mean() < 1


def test_predictor_keeps_the_speculative_work_on_the_realised_path(cuda, pkg):
    """Regression guard for the round-2 predictor: with the quadratic-in-1/Dc fit of the sum of squares,
    nearly every node of a chain's 32-lane tree lies on the path the chain then takes.  The balanced trees of round 1
    executed 6.2 solves per decided one at 32 lanes (3.6 at 16); the gate leaves a wide margin."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    c = 64
    q0 = np.random.default_rng(4).uniform(300.0, 4000.0, c)
    mc = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, nsamples=400, n_chains=c, verbose=False, seed=3)
    mc.sample(False)
    assert mc.stats["nsolves_executed"] < 1.5 * mc.stats["nsolves"], mc.stats
    assert 0.05 < float(np.mean(mc.acceptance_ratio)) < 0.95


def test_predictor_for_the_joint_posterior(cuda, pkg):
    """d = 3: the speculative kernel fits a full quadratic of the sum of squares in (1/a, b/a, 1/Dc) (ten
    coefficients, normal equations per group in shared memory).  Chains are those of the sequential kernel, and most
    of the tree lies on the realised path: acceptance-rate trees execute 3.6 (16 lanes) to 6 (32 lanes) solves per
    decided one."""
    g = load_golden("sse_grid.json")
    model = pkg.RateStateModel()
    c = 48
    rng = np.random.default_rng(6)
    q0 = np.stack([rng.uniform(0.0105, 0.0115, c), rng.uniform(0.0135, 0.0145, c), rng.uniform(1000.0, 1800.0, c)], axis=1)
    kw = dict(nsamples=240, n_chains=c, verbose=False, seed=12, param_names=("a", "b", "Dc"),
              bounds=[[0.005, 0.02], [0.005, 0.03], [0.0, 10000.0]])
    res = {}
    for depth in (1, 0):
        mc = pkg.MCMC(model, g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0, spec_depth=depth, **kw)
        out = mc.sample(False)
        res[depth] = (out, mc.std2.copy(), mc.accepts.copy(), dict(mc.stats))
    for a, b in zip(res[1][:3], res[0][:3]):
        assert np.array_equal(a, b)
    assert res[0][3]["nsolves"] == res[1][3]["nsolves"]
    assert res[1][3]["nsolves_executed"] == res[1][3]["nsolves"]
    assert res[0][3]["nsolves"] < res[0][3]["nsolves_executed"] < 2.6 * res[0][3]["nsolves"], res[0][3]
    assert 0.05 < res[0][2].mean() < 0.95
