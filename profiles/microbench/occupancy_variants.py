"""Wave quantisation of the one-thread-per-chain MCMC kernel (round 2).

65,536 chains (cfg 3) are 2,048 warps and 131,072 (cfg 5) are 4,096; at 168 registers 12 warps fit an SM = 1,776 on
148 SMs, i.e. 1.15 and 2.31 waves: the last, nearly empty wave runs latency-bound and costs almost a full one.
Builds capped at 128 registers (`__launch_bounds__(64, 7)` on rsf_mcmc_kernel, an experiment build) hold 16 warps per SM = 2,368: 0.86 and
1.73 waves.  This script times rsfm_run for both shapes with a given library and block size.

usage: python profiles/microbench/occupancy_variants.py <lib.so> <block_threads> [cfg3,cfg5,cfg2] [iters] [spec_depth]"""
import ctypes as C, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
if sys.argv[1] != "default":
    pkg._lib.LIB_PATH = os.path.abspath(sys.argv[1])
lib = pkg._lib.load()
import torch
import bench

block = int(sys.argv[2])
names = sys.argv[3].split(",") if len(sys.argv) > 3 else ["cfg3", "cfg5"]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 100
spec = int(sys.argv[5]) if len(sys.argv) > 5 else 0
for name in names:
    w = dict(bench.WORKLOADS[name]); w["name"] = name
    d, cpg = w["d"], w["chains"]
    model = pkg.RateStateModel(number_time_steps=w["n_out"], end_time=w["t_end"])
    model.a, model.b, model.Dc = w["truth"]
    np.random.seed(2024)
    data = model.evaluate()[2]
    model.a, model.b = 0.011, 0.014
    model.block_threads = block
    cfg = model.to_cfg()
    cfg.n_params, cfg.n_prior_len, cfg.spec_depth = d, 3, spec
    per = 200 if name == "cfg2" else 10
    for j in range(d):
        cfg.lo[j], cfg.hi[j] = w["lo"][j], w["hi"][j]
    q0 = bench.start_values(w, cpg, (0, cpg))
    q0_t = torch.from_numpy(np.ascontiguousarray(q0.reshape(cpg, d).T)).cuda()
    data_t = torch.from_numpy(data).cuda()
    h = lib.rsfm_create(C.byref(cfg), cpg, 20240, 0)
    pkg._lib.check(lib.rsfm_init(h, q0_t.data_ptr(), data_t.data_ptr(), None), "init")
    samples = torch.empty((per, d, cpg), dtype=torch.float64, device="cuda")
    for _ in range(3):
        pkg._lib.check(lib.rsfm_run(h, per, samples.data_ptr(), None, None, None, None), "run")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(max(1, iters // per)):
        pkg._lib.check(lib.rsfm_run(h, per, samples.data_ptr(), None, None, None, None), "run")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / max(1, iters // per)
    print(f"{os.path.basename(sys.argv[1])} block={block} {name} spec={spec} (depth {lib.rsfm_spec_depth(h)}): {ms:.2f} ms per {per}-iteration launch "
          f"(checksum {float(samples.sum()):.12e})", flush=True)
    tot = (C.c_uint64 * 9)()
    lib.rsfm_get_totals(h, tot, None)
    print("    totals", list(tot), flush=True)
    lib.rsfm_destroy(h)
