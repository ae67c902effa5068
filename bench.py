#!/usr/bin/env python
"""bench.py -- RSF-MCMC hot path: forward solves/s and ESS/s on N B200s, next to the reference CPU path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W     (reference CPU path on the host cores)

Workloads (BASELINE.json `configs`).  One STEP = `iters` Metropolis iterations of every chain; each
iteration of each chain is one forward solve unless the proposal leaves the prior box.  With the defaults the K
steps of a workload are its stated number of iterations, started from the start values (burn-in included):

    cfg3   configs[2], the largest single-GPU configuration = headline at N = 1: joint (a, b, Dc) posterior,
           65,536 chains, covariance pooled over all chains every 10 iterations after 100, 10 x 200 iterations
    cfg5   configs[4] = headline at N > 1: 131,072 chains per GPU (1,048,576 on 8), Dc posterior, proposal variance
           pooled over ALL chains of ALL ranks every 10 iterations -- partial sums all-gathered by NCCL on a side
           stream, moments / Cholesky / install on the device -- and split-R-hat / ESS all-reduced, all INSIDE
           the timed region; 5 x 100 iterations
    cfg2   configs[1]: 1,024 chains, Dc posterior, no adaptation (list priors as main.py), 10 x 200 iterations
    cfg4r  configs[3] reduced: velocity-step loading x10 every 1,000 s, stiff regime (Dc ~ 0.05), 16,384 chains,
           series of 20,000 points (2,000 s: one velocity step; the stated 100,000 points are a parity test and
           profiles/ record, 4-5 s of latency per solve), 4 x 1 iterations (a launch of one iteration has nothing
           to look ahead to and runs the one-thread-per-chain kernel; longer launches use two lanes per chain,
           +25 %: profiles/r2/pred/)
At N = 1 the line is the cfg3 record and carries cfg2 / cfg5 (one shard) / cfg4r as `sub_records`, each with its own
roofline (flops from ITS counters over ITS kernel time) and its own `traffic` (ncu capture of the same launch,
profiles/ncu_traffic.json).  `--workload X` runs one workload alone as the headline.

Per record:
  value      forward solves/s, whole job, inputs resident in HBM, CUDA-event time on the launching stream, max
             over ranks; also value_full_length_solves_only (solves stopped early by exact early rejection left out)
  e2e        the same metric through the public API (MCMC(...).sample()) with HOST buffers: data and start values
             copied host->device, samples / sigma^2 / accept flags copied device->host (overlapped with the
             iterations), setup solves included, wall clock
  roofline   FP64: algorithmic flops (35 n_rhs + 480 n_step of the solves that decided a proposal) /
             kernel time against the FP64 FMA peak measured live by rsfm_measure_fp64_peak; achieved HBM GB/s
             against MEASURED_PEAKS.json as evidence that memory is not the limiter
  cpu_baseline / --impl reference   the oracle's SciPy form (oracle/scipy_port.py: same scipy dop853 + Python RHS as
             the reference) on the host cores, a bounded sample of the SAME workload (stated in `sample`)
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "bayesian-markov-chain-monte-carlo_b200"

METRIC = "rsf_forward_solves_per_s"
UNIT = "solves/s"
FLOPS_PER_RHS, FLOPS_PER_STEP = 35.0, 480.0          # SURVEY.md 8d, source-level count

WORKLOADS = {
    "cfg2": dict(d=1, chains=1024, iters=200, steps=10, n_out=500, t_end=50.0, adapt=None, truth=(0.011, 0.014, 1325.0),
                 lo=[0.0], hi=[10000.0], loading="sine_decay",
                 text="cfg2: 1,024 independent chains per GPU, Dc-only posterior, N=500, list priors U(0,1e4) (no adaptation)"),
    "cfg3": dict(d=3, chains=65536, iters=200, steps=10, n_out=500, t_end=50.0, adapt="pooled", adapt_start=100,
                 truth=(0.011, 0.014, 1325.0), lo=[0.005, 0.005, 0.0], hi=[0.02, 0.03, 10000.0], loading="sine_decay",
                 text="cfg3: joint (a, b, Dc) posterior, 65,536 chains per GPU, N=500, proposal covariance pooled over all "
                      "chains every 10 iterations after 100 (device-side adaptive Metropolis)"),
    "cfg5": dict(d=1, chains=131072, iters=100, steps=5, n_out=500, t_end=50.0, adapt="pooled", adapt_start=100,
                 truth=(0.011, 0.014, 1325.0), lo=[0.0], hi=[10000.0], loading="sine_decay",
                 text="cfg5: 131,072 chains per GPU (1,048,576 on 8), Dc posterior, N=500, proposal variance pooled over all "
                      "chains of all ranks every 10 iterations after 100, split-R-hat / ESS pooled over ranks"),
    "cfg4r": dict(d=1, chains=16384, iters=1, steps=4, n_out=20000, t_end=2000.0, adapt=None, truth=(0.011, 0.014, 0.05),
                  lo=[0.01], hi=[1.0], loading="vstep",
                  text="cfg4 reduced: velocity steps x10 every 1,000 s, stiff regime Dc ~ 0.05, 16,384 chains, series of "
                       "20,000 points (cfg 4 states 100,000)"),
}


def start_values(w, total, first):
    """Deterministic start values of the global chain range [first[0], first[1])."""
    rng = np.random.default_rng(1)
    if w["loading"] == "vstep":
        q = rng.uniform(0.03, 0.08, size=total)
    elif w["d"] == 1:
        q = rng.uniform(200.0, 5000.0, size=total)
        q[0] = 1000.0                                   # chain 0 as the reference starts it (main.py)
    else:
        q = np.stack([rng.uniform(0.009, 0.013, total), rng.uniform(0.012, 0.016, total),
                      rng.uniform(800.0, 2500.0, total)], axis=1)
    return q[first[0]:first[1]]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle's SciPy form on the host cores
# ---------------------------------------------------------------------------------------------
def make_data_cpu(w):
    from oracle import scipy_port
    m = scipy_port.PortModel(number_time_steps=w["n_out"], end_time=w["t_end"])
    m.a, m.b, m.Dc = w["truth"]
    if w["loading"] == "vstep":
        m.loading, m.vstep_period, m.vstep_factor = "vstep", 1000.0, 10.0
    np.random.seed(2024)
    return m.evaluate()[2]


def cpu_port_sample(w, data, iters, processes):
    """`processes` independent chains of `iters` iterations of workload `w`; returns (solves, wall_s)."""
    from oracle import scipy_port
    q = start_values(w, max(processes, 2), (0, processes))
    if w["d"] == 3:
        r = scipy_port.run_chains_parallel(data, q, w["lo"], w["hi"], iters, seeds=range(100, 100 + processes),
                                           processes=processes, step_sd=[2.5e-4, 2.5e-4, 60.0])
    else:
        r = scipy_port.run_chains_parallel(data, q, w["lo"][0], w["hi"][0], iters, seeds=range(100, 100 + processes),
                                           processes=processes)
    return r["n_solves"], r["wall_s"]


def cpu_sample_text(w, cores, iters, steps=1):
    what = ("joint (a, b, Dc) random-walk Metropolis, fixed diagonal proposal" if w["d"] == 3
            else "Dc-only Metropolis as MCMC.py:494-527 (+3 setup solves per chain)")
    return (f"{steps} x ({cores} independent chains x {iters} iterations) of the {w['name']} workload (N={w['n_out']}), {what}; "
            "oracle/scipy_port.py = scipy ode('dop853') + Python RHS as the reference, one process per core")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                      # rank 0 alone runs the CPU arm
    w = pick_workload(args, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    if w["name"] == "cfg4r":
        raise SystemExit("the CPU arm of cfg4r would need ~1 h per step (1.5 s per RHS-callback solve interval); "
                         "use cfg2 / cfg3 / cfg5")
    cores = os.cpu_count() or 1
    data = make_data_cpu(w)
    iters = args.ref_iters
    for _ in range(args.warmup):
        cpu_port_sample(w, data, max(2, iters // 4), cores)
    solves, wall = 0, 0.0
    for _ in range(args.steps):
        s, t = cpu_port_sample(w, data, iters, cores)
        solves += s
        wall += t
    value = solves / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["text"] + f" -- CPU arm: a bounded sample of it, {cores} chains x {iters} iterations per step",
                   "name": w["name"], "chains": cores, "iters_per_step": iters, "n_out": w["n_out"], "n_params": w["d"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(w, cores, iters, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args, w):
    """Bounded CPU sample of workload `w` on the GPU box's host cores (rank 0, N = 1 only)."""
    cores = os.cpu_count() or 1
    data = make_data_cpu(w)
    iters = args.cpu_iters
    t0 = time.perf_counter()
    solves, wall = cpu_port_sample(w, data, iters, cores)
    out = {"value": solves / wall, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": cpu_sample_text(w, cores, iters), "wall_s": time.perf_counter() - t0}
    # context: the plain-C restatement (oracle/rsf_oracle.c) on all cores
    try:
        from oracle import oracle as orc
        dcs = np.random.default_rng(3).uniform(800.0, 2000.0, size=max(2000, 200 * cores))
        t1 = time.perf_counter()
        orc.forward_batch(orc.make_model(), dcs, data=make_data_cpu(WORKLOADS["cfg2"]), nthreads=cores)
        out["c_port"] = {"value": dcs.size / (time.perf_counter() - t1), "unit": UNIT, "cores": cores,
                         "sample": f"{dcs.size} forward solves + SSE (N=500), oracle/rsf_oracle.c, pthreads"}
    except Exception as ex:                                     # noqa: BLE001 - informational only
        out["c_port"] = {"error": str(ex)}
    return out


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.pkg = importlib.import_module(PKG)
        self.lib = self.pkg._lib.load()                 # raises if the CUDA extension is missing
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL prints its version banner on stdout when the communicator is created; the contract is
            # ONE JSON line on stdout, so file descriptor 1 points at stderr until the first collective is done
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize(self.dev)
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        if self.world != args.gpus and self.rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={self.world}; using WORLD_SIZE", file=sys.stderr)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)     # > 126 MB L2
        peak = C.c_double()
        self.pkg._lib.check(self.lib.rsfm_measure_fp64_peak(200.0, C.byref(peak)), "rsfm_measure_fp64_peak")
        self.fp64_peak = peak.value
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        self.traffic = {}
        try:
            self.traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except (OSError, ValueError):
            pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------
    def run(self, w, K, W, want_cpu=False, clocks=False):
        """One workload: device-resident arm (C ABI + the device-side adaptation pipeline), end-to-end arm
        (public API, host buffers), diagnostics; returns the record (rank 0) or None."""
        torch, dist, pkg, lib, dev, args = self.torch, self.dist, self.pkg, self.lib, self.dev, self.args
        world, rank = self.world, self.rank
        cpg, iters, d = w["chains"], w["iters"], w["d"]
        total_chains = cpg * world
        first = (rank * cpg, (rank + 1) * cpg)
        pooled = w["adapt"] == "pooled"
        interval = 10                                   # adapt_interval, MCMC.py:58

        # ---- synthetic data through the public forward model (GPU), same seed on every rank ----
        model = pkg.RateStateModel(number_time_steps=w["n_out"], end_time=w["t_end"])
        model.integ_mode = args.integ_mode
        model.chain_groups = args.chain_groups
        model.block_threads = args.block_threads
        model.round_packing = args.round_packing
        model.a, model.b, model.Dc = w["truth"]
        if w["loading"] == "vstep":
            model.loading, model.vstep_period, model.vstep_factor = "vstep", 1000.0, 10.0
        np.random.seed(2024)
        _, _, data = model.evaluate()
        model.a, model.b = 0.011, 0.014
        q0_host = start_values(w, total_chains, first)

        # ---- device-resident arm: everything in HBM before the timed region ----
        cfg = model.to_cfg()
        cfg.n_params, cfg.n_prior_len = d, 3
        cfg.adapt_mode = pkg._lib.ADAPT_POOLED if pooled else pkg._lib.ADAPT_NONE
        cfg.adapt_interval = interval
        cfg.spec_depth = args.spec_depth
        for j in range(d):
            cfg.lo[j], cfg.hi[j] = w["lo"][j], w["hi"][j]
        stream = pkg._lib.current_stream(torch, dev)
        data_t = torch.from_numpy(data).to(dev)
        q0_t = torch.from_numpy(np.ascontiguousarray(q0_host.reshape(cpg, d).T)).to(dev)
        PooledAdaptation = importlib.import_module(PKG + ".adaptation").PooledAdaptation

        def make_sampler():
            h = lib.rsfm_create(C.byref(cfg), cpg, C.c_uint64(args.seed), C.c_uint64(first[0]))
            if not h:
                pkg._lib.check(-1, "rsfm_create")
            pkg._lib.check(lib.rsfm_init(h, pkg._lib.ptr(q0_t), pkg._lib.ptr(data_t), stream), "rsfm_init")
            pool = PooledAdaptation(torch, lib, h, dev, d, total_chains, world, w.get("adapt_start", 0), stream) if pooled else None
            return h, pool

        samples = torch.empty((K * iters, d, cpg), dtype=torch.float64, device=dev)
        sigma2 = torch.empty((K * iters, cpg), dtype=torch.float64, device=dev)
        accept = torch.empty((K * iters, cpg), dtype=torch.uint8, device=dev)
        launches = [0, 0]                               # rsfm_run calls, adaptation intervals

        def run_iters(h, o, k):
            pkg._lib.check(lib.rsfm_run(h, k, pkg._lib.ptr(samples[o:]), pkg._lib.ptr(sigma2[o:]),
                                        pkg._lib.ptr(accept[o:]), None, stream), "rsfm_run")
            launches[0] += 1

        def step(h, pool, i):
            o = i * iters
            if pool is None:
                run_iters(h, o, iters)
                return
            for s0 in range(0, iters, interval):          # one launch per adaptation interval
                k = min(interval, iters - s0)
                pool.before_interval()
                run_iters(h, o + s0, k)
                pool.after_interval(o + s0 + k)
                launches[1] += 1

        def totals(h):
            out = (C.c_uint64 * 9)()
            pkg._lib.check(lib.rsfm_get_totals(h, out, stream), "rsfm_get_totals")
            return np.array(list(out), dtype=np.float64)

        # warm-up: W untimed steps of the same kernels on a throw-away sampler (same start values)
        hw, pw = make_sampler()
        for i in range(W):
            step(hw, pw, 0)
        if pw is not None:
            pw.finish()
        if W > 0:
            # the diagnostics kernels (and torch's reductions behind them) are part of the timed job: warm them too
            if samples.shape[0] >= 4:
                samples[iters:max(iters, 4)].zero_()
                importlib.import_module(PKG + ".diagnostics").chain_diagnostics(samples[:max(iters, 4)])
        torch.cuda.synchronize(dev)
        lib.rsfm_destroy(hw)
        # the timed job starts from the start values: K steps = the first K*iters iterations of every chain
        handle, pool = make_sampler()
        spec_g = int(lib.rsfm_spec_depth(handle))
        if (interval if pooled else iters) == 1:
            spec_g = 0                                  # a launch of one iteration runs the one-thread-per-chain kernel
        n_groups = int(lib.rsfm_chain_groups(handle))
        tot0 = totals(handle)
        launches[0] = launches[1] = 0
        sampler = ClockSampler(self.local) if (clocks and rank == 0) else None
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K + 1)]
        self.barrier()
        wall0 = time.perf_counter()
        for i in range(K):
            self.flush.zero_()                          # L2 flush between timed steps (not timed)
            ev[i][0].record()
            # chain groups (launches on the sampler's own streams): every group starts the step behind this point
            pkg._lib.check(lib.rsfm_join(handle, stream), "rsfm_join")
            step(handle, pool, i)
            ev[i][1].record()
        # convergence diagnostics of the second half of the draws: per-chain moments / ESS on the device, the sums
        # all-reduced over ranks (NCCL) -- part of the timed job
        ev[K][0].record()
        diag = importlib.import_module(PKG + ".diagnostics").chain_diagnostics(samples[(K * iters) // 2:])
        ev[K][1].record()
        self.barrier()
        wall = time.perf_counter() - wall0
        clock_info = sampler.stop() if sampler else None
        step_ms = [a.elapsed_time(b) for a, b in ev[:K]]
        diag_ms = ev[K][0].elapsed_time(ev[K][1])
        dev_s = (sum(step_ms) + diag_ms) * 1e-3
        kern_s = sum(step_ms) * 1e-3
        pool_stats = None
        if pool is not None:
            pool.finish()
            pool_stats = pool.stats
        tot = totals(handle) - tot0                     # (solves, nrhs, nstep, accepted, failed, early, executed, urhs, ustep)
        acc_rate = float(accept.float().mean().item())
        lib.rsfm_destroy(handle)
        # kernels launched in the timed region: the MCMC kernel once per chain group and rsfm_run; per adaptation
        # interval the partial sums (per group), moments / Cholesky (1) and install (per group); the diagnostics kernels
        n_launches = launches[0] * n_groups + launches[1] * (2 * n_groups + 1) + 3 * d

        # ---- end-to-end arm: public API, host buffers in, host results out, every step ----
        e2e = None
        if w.get("e2e", True):
            e2e_steps = max(1, args.e2e_steps)
            e2e_solves, e2e_wall, h2d, d2h = 0.0, 0.0, 0, 0
            pinned = torch.from_numpy(data).pin_memory()
            names = ("Dc",) if d == 1 else ("a", "b", "Dc")
            bounds = [[lo, hi] for lo, hi in zip(w["lo"], w["hi"])]
            self.barrier()
            out = mc = None
            for i in range(e2e_steps + 1):
                out = mc = None                         # the previous job's host arrays are consumed, not kept
                t0 = time.perf_counter()
                common = dict(nsamples=K * iters, verbose=False, seed=args.seed, device=dev, param_names=names, bounds=bounds,
                              adapt=w["adapt"], adapt_start=w.get("adapt_start", 100), adapt_interval=interval,
                              spec_depth=args.spec_depth)
                if world > 1 and pooled:                # chains sharded over the ranks, statistics pooled by NCCL
                    mc = pkg.MCMC(model, pinned.numpy(), w["truth"][2], ["Uniform", w["lo"][-1], w["hi"][-1]],
                                  start_values(w, total_chains, (0, total_chains)), n_chains=total_chains, shard=True, **common)
                else:
                    mc = pkg.MCMC(model, pinned.numpy(), w["truth"][2], ["Uniform", w["lo"][-1], w["hi"][-1]], q0_host,
                                  n_chains=cpg, chain_id0=first[0], **common)
                out = mc.sample(False)
                dt = time.perf_counter() - t0
                if i == 0:
                    continue                            # first call warms allocator / pinned staging buffers
                e2e_wall += dt
                e2e_solves += mc.stats["nsolves"]
                h2d = data.nbytes + q0_host.nbytes
                d2h = out.nbytes + mc.std2.nbytes + mc.accepts.nbytes
            self.barrier()
            e2e = (e2e_solves, e2e_wall, h2d, d2h, e2e_steps)
            out = mc = None

        # ---- reduce over ranks: times MAX, work SUM ----
        red_max = torch.tensor([dev_s, wall, e2e[1] if e2e else 0.0, kern_s], dtype=torch.float64, device=dev)
        red_sum = torch.tensor(list(tot) + [e2e[0] if e2e else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(red_max, op=dist.ReduceOp.MAX)
            dist.all_reduce(red_sum, op=dist.ReduceOp.SUM)
        dev_s_max, wall_max, e2e_wall_max, kern_s_max = red_max.tolist()
        solves, nrhs, nstep, n_accepted, n_failed, n_early, n_exec, urhs, ustep, e2e_solves_all = red_sum.tolist()
        if rank != 0:
            return None

        peak_tf = self.fp64_peak / 1e12
        hbm_peak = self.peaks.get("hbm_gbs", 6650.0)
        # algorithmic flops of the solves that decided a proposal, this rank (speculative work that was thrown
        # away is NOT counted as achieved; it is reported separately as `executed`)
        achieved_tf = (FLOPS_PER_RHS * tot[7] + FLOPS_PER_STEP * tot[8]) / kern_s / 1e12
        executed_tf = (FLOPS_PER_RHS * tot[1] + FLOPS_PER_STEP * tot[2]) / kern_s / 1e12
        # algorithmic bytes per kernel launch: state in/out 64 B/chain, per iteration 8 d B sample +
        # 8 B sigma^2 + 1 B flag per chain, and the 8 N B series once per block
        per_launch_iters = interval if pooled else iters
        n_kernel_launches = K * (iters // per_launch_iters if pooled else 1)
        threads = cpg << spec_g if spec_g >= 1 else cpg
        block = 32 if (w["loading"] == "vstep" or threads <= 148 * 32) else 64 if threads <= 148 * 128 else 128
        if spec_g >= 1:
            block = 32                                   # the speculative kernel runs one-warp blocks
        nblocks = (threads + block - 1) // block
        alg_bytes = cpg * 64.0 + per_launch_iters * cpg * (8.0 * d + 9.0) + nblocks * 8.0 * w["n_out"]
        launch_s = kern_s / n_kernel_launches
        if spec_g >= 1:
            kernel = f"rsf_mcmc_spec_kernel<{d},false,{'true' if w['loading'] == 'vstep' else 'false'}> (speculation depth {spec_g}: {1 << spec_g} lanes per chain)"
        else:
            kernel = f"rsf_mcmc_kernel<{d},false,{'true' if w['loading'] == 'vstep' else 'false'}> (one thread per chain)"
            if n_groups > 1:
                kernel += (f"; each {per_launch_iters}-iteration launch is {n_groups} launches of {cpg // n_groups} chains on the "
                           "sampler's own streams (chain groups), timed as one")
        tr = self.traffic.get(w["name"], {})
        traffic = tr.get("dram_bytes_per_launch") if (tr.get("chains_per_gpu") == cpg and tr.get("iters_per_launch") == per_launch_iters) else None
        rec = {
            "metric": METRIC, "value": solves / dev_s_max, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_s_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["text"] + f"; {iters} Metropolis iterations per step, {K} steps from the start values",
                       "name": w["name"], "chains_per_gpu": cpg, "chains_total": total_chains, "iters_per_step": iters,
                       "n_out": w["n_out"], "n_params": d, "integ_mode": args.integ_mode, "speculation_depth": spec_g,
                       "chain_groups": n_groups,
                       "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": (f"chains sharded over {world} GPU(s); one NCCL all-gather of the pooled partial sums per "
                                       f"{interval} iterations on a side stream (overlapped with the next kernel, adaptation lags one "
                                       "interval), all-reduces of the R-hat / ESS sums at the end -- all inside the timed region"
                                       if pooled and world > 1 else
                                       f"chains sharded over {world} GPU(s); " + ("pooled adaptation on the device, " if pooled else "no data-path collective, ") +
                                       ("all-reduces of the R-hat / ESS sums inside the timed region" if world > 1 else "single rank"))},
            # transparency: proposals whose solve ran to the end of the series (early-rejected ones excluded)
            "value_full_length_solves_only": (solves - n_early) / dev_s_max,
            "chain_iterations_per_s": total_chains * K * iters / dev_s_max,
            "ess_per_s": min(diag["ess"]) / dev_s_max,
            "ess": {"total_min_over_params": min(diag["ess"]), "total": diag["ess"], "per_chain": diag["ess_per_chain_mean"],
                    "rhat": diag["rhat"], "draws_per_chain": int(diag["n"]), "posterior_mean": diag["mean"],
                    "posterior_sd": diag["sd"], "accept_rate": acc_rate},
            "work": {"forward_solves": solves, "of_which_stopped_early": n_early,
                     "solves_executed_incl_speculative": n_exec, "rhs_evals": nrhs, "ode_steps": nstep,
                     "rhs_evals_deciding": urhs, "ode_steps_deciding": ustep, "failed_chains": n_failed,
                     "wall_s_timed_region": wall_max, "kernel_s": kern_s_max, "diagnostics_ms": diag_ms},
            "gpu_launches": n_launches,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": traffic,
                         "traffic_source": tr.get("source") if traffic is not None else None,
                         "executed_incl_speculative": executed_tf, "executed_frac": executed_tf / peak_tf,
                         "kernel": kernel, "launches": n_kernel_launches, "avg_launch_ms": 1e3 * launch_s,
                         "peak_source": "rsfm_measure_fp64_peak (DFMA chains, live; MEASURED_PEAKS.json has no FP64 entry)",
                         "flops_convention": "35 per RHS + 480 per DOP853 step (SURVEY.md 8d), deciding solves only",
                         "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / launch_s / 1e9,
                                 "peak_gbs": hbm_peak, "frac": alg_bytes / launch_s / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if self.peaks else "fallback"}},
        }
        if pool_stats is not None:
            rec["adaptation"] = {"n_adaptations": pool_stats["n_adaptations"], "n_intervals": pool_stats["n_intervals"],
                                 "rows_gathered_per_interval": pool_stats["pool_rows_gathered"],
                                 "allgather_ms_total_side_stream": pool_stats["collective_ms_on_side_stream"],
                                 "where": "partial sums, moments, Cholesky and install on the device; no host synchronisation"}
        if e2e is not None:
            rec["e2e"] = {"value": e2e_solves_all / e2e_wall_max, "unit": UNIT, "h2d_bytes_per_step": int(e2e[2]),
                          "d2h_bytes_per_step": int(e2e[3]), "steps": e2e[4], "wall_s_per_job": e2e_wall_max / e2e[4],
                          "wall_over_device_time": (e2e_wall_max / e2e[4]) / dev_s_max,
                          "call": "one MCMC(model, data_host, ...).sample() per e2e step = the whole timed job: data / start "
                                  f"values H2D, setup solves, {K * iters} iterations, post-burn-in samples / sigma2 / accepts D2H "
                                  "(overlapped with the iterations)"}
        if clock_info is not None:
            rec["clocks"] = clock_info
        if want_cpu:
            rec["cpu_baseline"] = cpu_baseline(args, w)
        return rec


def _json_safe(o):
    """NaN / inf (e.g. R-hat of a two-draw record) as null: the line must parse under strict JSON."""
    if isinstance(o, dict):
        return {k: _json_safe(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_json_safe(v) for v in o]
    if isinstance(o, float) and not np.isfinite(o):
        return None
    return o


def pick_workload(args, world):
    name = args.workload or ("cfg3" if world == 1 else "cfg5")
    w = dict(WORKLOADS[name])
    w["name"] = name
    if args.chains:
        w["chains"] = args.chains
    if args.iters:
        w["iters"] = args.iters
    return w


def run_b200(args):
    b = Bench(args)
    w = pick_workload(args, b.world)
    K = args.steps if args.steps else w["steps"]
    W = max(3, args.warmup)                             # timing rule: W >= 3
    want_cpu = b.world == 1 and not args.no_cpu_baseline and w["name"] != "cfg4r"
    rec = b.run(w, K, W, want_cpu=want_cpu, clocks=True)
    subs = {}
    if b.world == 1 and not args.workload and not args.no_sub_records:
        for name in ("cfg2", "cfg5", "cfg4r"):
            sw = dict(WORKLOADS[name])
            sw["name"] = name
            if name == "cfg4r":
                sw["e2e"] = False
            r = b.run(sw, sw["steps"], 3)
            if r is not None:
                for k in ("n_gpus", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "metric", "unit"):
                    r.pop(k, None)
                subs[name] = r
    if b.rank == 0:
        if subs:
            rec["sub_records"] = subs
        print(json.dumps(_json_safe(rec)), flush=True)
    if b.world > 1:
        b.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (0 = the workload's own count)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=None,
                    help="default: cfg3 at N = 1 (with cfg2 / cfg5 / cfg4r as sub-records), cfg5 at N > 1")
    ap.add_argument("--chains", type=int, default=0, help="chains per GPU (0 = the workload's own)")
    ap.add_argument("--iters", type=int, default=0, help="Metropolis iterations per step (0 = the workload's own)")
    ap.add_argument("--integ-mode", choices=["parity", "carry"], default="parity")
    ap.add_argument("--chain-groups", type=int, default=0,
                    help="pooled adaptation: launches per interval on the sampler's own streams (0 auto, 1 off); "
                         "never affects results")
    ap.add_argument("--round-packing", type=int, default=0, help="d = 3 kernel: 0 auto (on), 1 off; never affects results")
    ap.add_argument("--block-threads", type=int, default=0, help="threads per block of the one-thread-per-chain kernels (0 auto)")
    ap.add_argument("--spec-depth", type=int, default=0,
                    help="few-chain workloads: lanes per chain of the speculative kernel = 2^depth (0 auto, 1 off, 2..5); never affects results")
    ap.add_argument("--seed", type=int, default=20240)
    ap.add_argument("--e2e-steps", type=int, default=2, help="timed public-API calls (each = the whole job)")
    ap.add_argument("--cpu-iters", type=int, default=12, help="iterations per chain in the cpu_baseline sample")
    ap.add_argument("--ref-iters", type=int, default=12, help="iterations per chain per step, --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if not args.steps:
            args.steps = 1
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
