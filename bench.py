#!/usr/bin/env python
"""bench.py -- RSF-MCMC hot path: forward solves/s and ESS/s on N B200s, next to the reference CPU path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference --steps K --warmup W     (reference CPU path on the host cores)

Workload (BASELINE.json configs[1], SURVEY.md 8d cfg 2): 1,024 independent chains per GPU, Dc-only
posterior, N = 500 output points over T = [0, 50], synthetic data = acc(Dc_true = 1325) + |acc| N(0,1)
with np.random.seed(2024), prior U(0, 10000) (list form: no adaptation, as main.py), chain 0 starts at
1000 and the others at U(200, 5000).  One STEP = `--iters` (default 200) Metropolis iterations of every
chain = one launch of rsf_mcmc_kernel; each iteration of each chain is one forward solve (unless the
proposal is out of bounds).  The timed job starts at the start values, so the default ten steps are
exactly cfg 2's nsamples = 2,000 (burn-in phase included); warm-up steps run on a throw-away sampler.

Reported on one JSON line:
  value      forward solves/s, whole job, inputs resident in HBM, CUDA-event time, max over ranks
  e2e        the same metric through the public API (MCMC(...).sample()) with HOST buffers: data and
             start values copied host->device, samples / sigma^2 / accept flags copied device->host,
             setup solves included, wall clock
  roofline   FP64: algorithmic flops (35 n_rhs + 480 n_step) / kernel time against the
             FP64 FMA peak measured live by rsfm_measure_fp64_peak; plus achieved HBM GB/s against
             MEASURED_PEAKS.json as evidence that memory is not the limiter
  cpu_baseline  the oracle's SciPy form (oracle/scipy_port.py: same scipy dop853 + Python RHS as the
             reference) timed on the host cores on a bounded sample
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
DC_TRUE, LO, HI, N_OUT = 1325.0, 0.0, 10000.0, 500
FLOPS_PER_RHS, FLOPS_PER_STEP = 35.0, 480.0          # SURVEY.md 8d, source-level count


def start_values(total, first):
    """Chain 0 starts at 1000 (reference-compatible), the others at U(200, 5000); deterministic."""
    rng = np.random.default_rng(1)
    q = rng.uniform(200.0, 5000.0, size=total)
    q[0] = 1000.0
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
def make_data_cpu():
    from oracle import scipy_port
    m = scipy_port.PortModel(number_time_steps=N_OUT)
    m.Dc = DC_TRUE
    np.random.seed(2024)
    return m.evaluate()[2]


def cpu_port_sample(data, iters, processes):
    """`processes` independent chains of `iters` iterations each; returns (solves, wall_s, chains)."""
    from oracle import scipy_port
    q = start_values(max(processes, 1), (0, max(processes, 1)))
    r = scipy_port.run_chains_parallel(data, q, LO, HI, iters, seeds=range(100, 100 + processes), processes=processes)
    return r["n_solves"], r["wall_s"], r["chains"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                      # rank 0 alone runs the CPU arm
    cores = os.cpu_count() or 1
    data = make_data_cpu()
    iters = args.ref_iters
    for _ in range(args.warmup):
        cpu_port_sample(data, max(2, iters // 4), cores)
    solves, wall, ess = 0, 0.0, 0.0
    for _ in range(args.steps):
        s, w, chains = cpu_port_sample(data, iters, cores)
        solves += s
        wall += w
    value = solves / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2 sample: RSF-MCMC Dc-only posterior, N=500, list priors U(0,1e4); "
                               f"{cores} independent chains x {iters} iterations per step on the host cores",
                   "chains": cores, "iters_per_step": iters, "n_out": N_OUT},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x ({cores} chains x {iters} Metropolis iterations), "
                                   "oracle/scipy_port.py = scipy ode('dop853') + Python RHS, one process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # the chain law is the reference's, so ESS per draw is a property of the algorithm: 23.7 effective
        # samples in 251 post-burn-in draws for the unmodified reference at Dc_true = 1325
        "ess_per_s_estimate": value * 23.7 / 251.0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    lib = pkg._lib.load()                           # raises if the CUDA extension is missing
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created; the contract is
        # ONE JSON line on stdout, so file descriptor 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if world != args.gpus:
        if rank == 0:
            print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    cpg, iters, K, W = args.chains, args.iters, args.steps, args.warmup
    total_chains = cpg * world
    first = (rank * cpg, (rank + 1) * cpg)

    # ---- synthetic data through the public forward model (GPU), same seed on every rank ----
    model = pkg.RateStateModel(number_time_steps=N_OUT)
    model.integ_mode = args.integ_mode
    model.Dc = DC_TRUE
    np.random.seed(2024)
    _, _, data = model.evaluate()
    q0_host = start_values(total_chains, first)

    # ---- device-resident arm: C ABI directly, everything in HBM before the timed region ----
    cfg = model.to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.adapt_mode = 1, 3, pkg._lib.ADAPT_NONE
    cfg.lo[0], cfg.hi[0] = LO, HI
    stream = pkg._lib.current_stream(torch, dev)
    data_t = torch.from_numpy(data).to(dev)
    q0_t = torch.from_numpy(q0_host.reshape(1, -1).copy()).to(dev)
    def make_sampler():
        h = lib.rsfm_create(C.byref(cfg), cpg, C.c_uint64(args.seed), C.c_uint64(first[0]))
        if not h:
            pkg._lib.check(-1, "rsfm_create")
        pkg._lib.check(lib.rsfm_init(h, pkg._lib.ptr(q0_t), pkg._lib.ptr(data_t), stream), "rsfm_init")
        return h

    samples = torch.empty((K * iters, 1, cpg), dtype=torch.float64, device=dev)
    sigma2 = torch.empty((K * iters, cpg), dtype=torch.float64, device=dev)
    accept = torch.empty((K * iters, cpg), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2

    def step(h, i):
        o = i * iters
        pkg._lib.check(lib.rsfm_run(h, iters, pkg._lib.ptr(samples[o:]), pkg._lib.ptr(sigma2[o:]),
                                    pkg._lib.ptr(accept[o:]), None, stream), "rsfm_run")

    def totals(h):
        out = (C.c_uint64 * 9)()
        pkg._lib.check(lib.rsfm_get_totals(h, out, stream), "rsfm_get_totals")
        return np.array(list(out), dtype=np.float64)

    # warm-up: W untimed steps of the same kernel on a throw-away sampler (same start values)
    hw = make_sampler()
    for _ in range(W):
        step(hw, 0)
    torch.cuda.synchronize(dev)
    lib.rsfm_destroy(hw)
    # the timed job starts from the start values: K steps = the first K*iters iterations of every
    # chain (K = 10: cfg 2's nsamples = 2,000, burn-in phase included), data and state resident in HBM
    handle = make_sampler()
    spec_g = int(lib.rsfm_spec_depth(handle))
    tot0 = totals(handle)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()                               # L2 flush between timed steps (not timed)
        ev[i][0].record()
        step(handle, i)
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    clock_info = clocks.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_s = sum(step_ms) * 1e-3
    tot = totals(handle) - tot0                     # (solves, nrhs, nstep, accepted, failed, early)

    # ---- ESS of the second half of the timed draws (R-hat alongside), pooled over ranks ----
    diag = importlib.import_module(PKG + ".diagnostics").chain_diagnostics(samples[(K * iters) // 2:])
    acc_rate = float(accept.float().mean().item())
    lib.rsfm_destroy(handle)

    # ---- end-to-end arm: public API, host buffers in, host results out, every step ----
    e2e_steps = max(1, args.e2e_steps)
    e2e_solves, e2e_wall, h2d, d2h = 0.0, 0.0, 0, 0
    pinned = torch.from_numpy(data).pin_memory()
    barrier()
    out = mc = None
    for i in range(e2e_steps + 1):
        out = mc = None                             # the previous job's host arrays are consumed, not kept
        t0 = time.perf_counter()
        mc = pkg.MCMC(model, pinned.numpy(), DC_TRUE, ["Uniform", LO, HI], q0_host, nsamples=K * iters,
                      n_chains=cpg, verbose=False, seed=args.seed, device=dev, chain_id0=first[0])
        out = mc.sample(False)
        dt = time.perf_counter() - t0
        if i == 0:
            continue                                # first call warms allocator / module state
        e2e_wall += dt
        e2e_solves += mc.stats["nsolves"]
        h2d = data.nbytes + q0_host.nbytes
        d2h = out.nbytes + mc.std2.nbytes + mc.accepts.nbytes
    barrier()

    # ---- FP64 peak (roofline denominator), measured live on this GPU ----
    peak = C.c_double()
    pkg._lib.check(lib.rsfm_measure_fp64_peak(200.0, C.byref(peak)), "rsfm_measure_fp64_peak")

    # ---- reduce over ranks: times MAX, work SUM ----
    red_max = torch.tensor([dev_s, wall, e2e_wall], dtype=torch.float64, device=dev)
    red_sum = torch.tensor(list(tot) + [e2e_solves], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(red_sum, op=dist.ReduceOp.SUM)
    dev_s_max, wall_max, e2e_wall_max = red_max.tolist()
    solves, nrhs, nstep, n_accepted, n_failed, n_early, n_exec, urhs, ustep, e2e_solves_all = red_sum.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic flops of the solves that decided a proposal (speculative work that was thrown
        # away is NOT counted as achieved; it is reported separately as `executed`)
        flops_rank0 = FLOPS_PER_RHS * tot[7] + FLOPS_PER_STEP * tot[8]
        achieved_tf = flops_rank0 / dev_s / 1e12
        executed_tf = (FLOPS_PER_RHS * tot[1] + FLOPS_PER_STEP * tot[2]) / dev_s / 1e12
        # algorithmic bytes per launch: state in/out 64 B/chain, per iteration 8 B sample +
        # 8 B sigma^2 + 1 B flag per chain, and the 8 N B series once per block
        nblocks = (cpg + 31) // 32 if cpg <= 148 * 32 else (cpg + 63) // 64 if cpg <= 148 * 128 else (cpg + 127) // 128
        alg_bytes = cpg * 64.0 + iters * cpg * 17.0 + nblocks * 8.0 * N_OUT
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        cpu = cpu_baseline(args) if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": METRIC, "value": solves / dev_s_max, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dev_s_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfg2: {cpg} independent chains per GPU, Dc-only posterior, N={N_OUT}, "
                                   f"list priors U(0,1e4), {iters} Metropolis iterations per step",
                       "chains_per_gpu": cpg, "chains_total": total_chains, "iters_per_step": iters, "n_out": N_OUT,
                       "integ_mode": args.integ_mode, "speculation_depth": spec_g, "l2": "flushed between timed steps (256 MiB write)",
                       "parallelism": f"chains sharded over {world} GPU(s), no data-path collective"},
            # transparency: proposals whose solve ran to the end of the series (early-rejected ones excluded)
            "value_full_length_solves_only": (solves - n_early) / dev_s_max,
            "ess_per_s": diag["ess"][0] / dev_s_max,
            "ess": {"total": diag["ess"][0], "per_chain": diag["ess_per_chain_mean"][0], "rhat": diag["rhat"][0],
                    "draws_per_chain": int(diag["n"]), "posterior_mean": diag["mean"][0], "posterior_sd": diag["sd"][0],
                    "accept_rate": acc_rate},
            "work": {"forward_solves": solves, "of_which_stopped_early": n_early,
                     "solves_executed_incl_speculative": n_exec, "rhs_evals": nrhs,
                     "ode_steps": nstep, "rhs_evals_deciding": urhs, "ode_steps_deciding": ustep,
                     "failed_chains": n_failed,
                     "wall_s_timed_region": wall_max},
            "e2e": {"value": e2e_solves_all / e2e_wall_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "call": "one MCMC(model, data_host, ...).sample() per e2e step = the whole timed job: data/start "
                            f"values H2D, setup solves, {K * iters} iterations, post-burn-in samples/sigma2/accepts D2H"},
            "gpu_launches": K,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak.value / 1e12, "unit": "TFLOP/s",
                         "frac": achieved_tf / (peak.value / 1e12), "traffic": traffic,
                         "executed_incl_speculative": executed_tf,
                         "executed_frac": executed_tf / (peak.value / 1e12),
                         "kernel": (f"rsf_mcmc_spec_kernel<1,false> (speculation depth {spec_g}: {1 << spec_g} lanes per chain)"
                                    if spec_g >= 2 else "rsf_mcmc_kernel<1,false> (one thread per chain)"), "peak_source": "rsfm_measure_fp64_peak (DFMA chains, live)",
                         "flops_convention": "35 per RHS + 480 per DOP853 step (SURVEY.md 8d)",
                         "hbm": {"achieved_gbs": alg_bytes / (dev_s / K) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": alg_bytes / (dev_s / K) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "clocks": clock_info,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    """Bounded CPU sample on the GPU box's host cores (rank 0, N = 1 only)."""
    cores = os.cpu_count() or 1
    data = make_data_cpu()
    iters = args.cpu_iters
    t0 = time.perf_counter()
    solves, wall, chains = cpu_port_sample(data, iters, cores)
    out = {"value": solves / wall, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{cores} independent chains x {iters} Metropolis iterations (+3 setup solves each), "
                     "oracle/scipy_port.py = scipy ode('dop853') + Python RHS as the reference, one process per core",
           "wall_s": time.perf_counter() - t0}
    # context: the plain-C restatement (oracle/rsf_oracle.c) on all cores
    try:
        from oracle import oracle as orc
        dcs = np.random.default_rng(3).uniform(800.0, 2000.0, size=max(2000, 200 * cores))
        t1 = time.perf_counter()
        orc.forward_batch(orc.make_model(), dcs, data=data, nthreads=cores)
        out["c_port"] = {"value": dcs.size / (time.perf_counter() - t1), "unit": UNIT, "cores": cores,
                         "sample": f"{dcs.size} forward solves + SSE, oracle/rsf_oracle.c, pthreads"}
    except Exception as ex:                                     # noqa: BLE001 - informational only
        out["c_port"] = {"error": str(ex)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU")
    ap.add_argument("--iters", type=int, default=200, help="Metropolis iterations per step")
    ap.add_argument("--integ-mode", choices=["parity", "carry"], default="parity")
    ap.add_argument("--seed", type=int, default=20240)
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed public-API calls (each = the whole job)")
    ap.add_argument("--cpu-iters", type=int, default=12, help="iterations per chain in the cpu_baseline sample")
    ap.add_argument("--ref-iters", type=int, default=12, help="iterations per chain per step, --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                              # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
