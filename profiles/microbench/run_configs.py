"""BASELINE.json configs 3-5 at reduced iteration counts (capability + throughput data points, not bench lines).

usage: python profiles/microbench/run_configs.py cfg3|cfg4|cfg5 [iters]
"""
import importlib, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
which = sys.argv[1]
rng = np.random.default_rng(0)
SHARD = int(os.environ.get("WORLD_SIZE", "1")) > 1        # under torchrun: chains sharded over ranks, NCCL all-reduces
if SHARD:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    sys.stdout.flush(); _saved = os.dup(1); os.dup2(2, 1)      # keep the NCCL banner off stdout
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    dist.barrier(); os.dup2(_saved, 1); os.close(_saved)


def report(name, mc, out, wall, extra=None):
    d = mc.diagnostics()
    line = {"config": name, "wall_s": wall, "solves": mc.stats["nsolves"], "solves_per_s": mc.stats["nsolves"] / wall,
            "executed_solves": mc.stats["nsolves_executed"], "accept_rate": float(np.mean(mc.acceptance_ratio)),
            "rhat": d["rhat"], "ess_total": d["ess"], "ess_per_s": [e / wall for e in d["ess"]],
            "post_mean": d["mean"], "post_sd": d["sd"], "failed_chains": mc.stats["failed_chains"]}
    if extra:
        line.update(extra)
    if SHARD:
        line["rank"] = dist.get_rank(); line["world"] = dist.get_world_size()
        line["chains_local"] = mc.stats["n_chains_local"]
    print(json.dumps(line), flush=True)


if which == "cfg1":
    # the reference's own default (main.py:50-56): 5 Dc values, one chain each, 500 samples, through the RSF facade
    import contextlib, io, tempfile
    os.chdir(tempfile.mkdtemp())
    problem = pkg.RSF(number_slip_values=5, lowest_slip_value=100.0, largest_slip_value=5000.0, qstart=1000.0,
                      qpriors=["Uniform", 0.0, 10000.0])
    problem.model = pkg.RateStateModel(number_time_steps=500)
    np.random.seed(2024)
    t0 = time.perf_counter(); problem.data = problem.generate_time_series(); t_gen = time.perf_counter() - t0
    problem.format = "json"
    problem.mcmc_kwargs = {"seed": 2024}
    with contextlib.redirect_stdout(io.StringIO()):
        elapsed = problem.inference(nsamples=500)
    post = {float(dc): (float(r["samples"].mean()), float(r["samples"].std()), float(r["acceptance_ratio"][0]))
            for dc, r in problem.results.items()}
    print(json.dumps({"config": "cfg1: main.py defaults via RSF facade (5 x 1 chain x 500 samples)",
                      "generate_time_series_s": t_gen, "inference_s": elapsed, "posterior_mean_sd_accept": post}))
elif which == "cfg3":
    # joint (a, b, Dc), pooled adaptive covariance, 65,536 chains
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    c = int(os.environ.get("CHAINS", "65536"))
    m = pkg.RateStateModel()
    m.a, m.b, m.Dc = 0.011, 0.014, 1325.0
    np.random.seed(2024)
    _, _, data = m.evaluate()
    q0 = np.stack([rng.uniform(0.009, 0.013, c), rng.uniform(0.012, 0.016, c), rng.uniform(800.0, 2500.0, c)], axis=1)
    mc = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0, nsamples=iters, n_chains=c, verbose=False, seed=3,
                  param_names=("a", "b", "Dc"), bounds=[[0.005, 0.02], [0.005, 0.03], [0.0, 1e4]], adapt="pooled",
                  adapt_start=100, shard=SHARD)
    t0 = time.perf_counter(); out = mc.sample(False); wall = time.perf_counter() - t0
    report("cfg3: joint (a,b,Dc), pooled AM, %d chains, %d iters" % (c, iters), mc, out, wall,
           {"n_adaptations": len(mc.adapt_history)})
elif which == "cfg4":
    # long series N = 1e5, VSTEP loading, stiff regime, 16,384 chains
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    c = int(os.environ.get("CHAINS", "16384"))
    n = int(os.environ.get("NOUT", "100000"))
    m = pkg.RateStateModel(number_time_steps=n, end_time=n * 0.1)
    m.loading, m.vstep_period, m.vstep_factor = "vstep", 1000.0, 10.0
    m.Dc = 0.05
    np.random.seed(2024)
    t0 = time.perf_counter(); _, acc, data = m.evaluate(); t_one = time.perf_counter() - t0
    q0 = rng.uniform(0.03, 0.08, c)
    mc = pkg.MCMC(m, data, 0.05, ["Uniform", 0.01, 1.0], q0, nsamples=iters, n_chains=c, verbose=False, seed=4)
    t0 = time.perf_counter(); out = mc.sample(False); wall = time.perf_counter() - t0
    report("cfg4: N=%d VSTEP stiff, %d chains, %d iters" % (n, c, iters), mc, out, wall,
           {"single_solve_s": t_one, "rhs_per_solve": mc.stats["nrhs"] / max(mc.stats["nsolves_executed"], 1)})
elif which == "cfg5":
    # 131,072 chains per GPU (1,048,576 over 8 GPUs under torchrun), Dc-only, proposal variance pooled over ALL
    # chains of ALL ranks by NCCL all-reduces of the sufficient statistics, R-hat / ESS pooled the same way
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    c = int(os.environ.get("CHAINS", "131072")) * world
    m = pkg.RateStateModel(); m.Dc = 1325.0
    np.random.seed(2024)
    _, _, data = m.evaluate()
    q0 = rng.uniform(200.0, 5000.0, c); q0[0] = 1000.0
    mc = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0, nsamples=iters, n_chains=c, verbose=False, seed=5,
                  adapt="pooled", adapt_start=50, shard=SHARD)
    warm = pkg.MCMC(m, data, 1325.0, ["Uniform", 0.0, 1e4], q0, nsamples=20, n_chains=c, verbose=False, seed=6,
                    adapt="pooled", adapt_start=10, shard=SHARD)
    warm.sample(False); del warm                      # module load, NCCL communicator, pinned staging buffers
    if SHARD: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); out = mc.sample(False); wall = time.perf_counter() - t0
    tot = torch.tensor([float(mc.stats["nsolves"]), float(mc.stats["nsolves_executed"])], dtype=torch.float64, device="cuda")
    wmax = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if SHARD:
        dist.all_reduce(tot); dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    report("cfg5: %d chains over %d GPU(s), %d iters, pooled adaptation" % (c, world, iters), mc, out, wall,
           {"n_adaptations": len(mc.adapt_history),
            "proposal_sd_first_last": [float(np.sqrt(mc.adapt_history[0][1][0])), float(np.sqrt(mc.adapt_history[-1][1][0]))] if mc.adapt_history else [],
            "all_ranks": {"chains": c, "solves": tot[0].item(), "wall_max_s": wmax.item(), "solves_per_s": tot[0].item() / wmax.item(),
                          "chain_iterations_per_s": c * iters / wmax.item()}})
