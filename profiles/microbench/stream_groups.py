"""Does splitting the chains of one GPU into G groups on G streams hide the tail of each launch?

Every rsfm_run launch of the pooled-adaptation pipeline is 10 iterations long; at 65,536 chains (cfg 3 shape) the 512
blocks are 1.15 waves and ncu shows the SMs idle 18 % of the launch.  Groups on their own streams have no barrier in
common, so the tail of one group's launch overlaps the other groups' work.  Same chains either way (global chain ids).

usage: python profiles/microbench/stream_groups.py cfg3|cfg5 G [iters]"""
import ctypes as C, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
lib = pkg._lib.load()
import torch
import bench

name, G = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
w = dict(bench.WORKLOADS[name]); w["name"] = name
d, cpg = w["d"], w["chains"]
model = pkg.RateStateModel(number_time_steps=w["n_out"], end_time=w["t_end"])
model.a, model.b, model.Dc = w["truth"]
np.random.seed(2024)
data = model.evaluate()[2]
model.a, model.b = 0.011, 0.014
cfg = model.to_cfg()
cfg.n_params, cfg.n_prior_len = d, 3
for j in range(d):
    cfg.lo[j], cfg.hi[j] = w["lo"][j], w["hi"][j]
q0 = bench.start_values(w, cpg, (0, cpg)).reshape(cpg, d)
data_t = torch.from_numpy(data).cuda()
per = cpg // G
streams = [torch.cuda.Stream() for _ in range(G)]
handles, outs = [], []
for g in range(G):
    q0_t = torch.from_numpy(np.ascontiguousarray(q0[g * per:(g + 1) * per].T)).cuda()
    h = lib.rsfm_create(C.byref(cfg), per, 20240, g * per)
    pkg._lib.check(lib.rsfm_init(h, q0_t.data_ptr(), data_t.data_ptr(), None), "init")
    handles.append(h)
    outs.append(torch.empty((10, d, per), dtype=torch.float64, device="cuda"))
torch.cuda.synchronize()


def sweep(n_launch):
    for _ in range(n_launch):
        for g in range(G):
            pkg._lib.check(lib.rsfm_run(handles[g], 10, outs[g].data_ptr(), None, None, None, streams[g].cuda_stream), "run")


sweep(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in streams:
    s.wait_event(e0)
sweep(iters // 10)
for s in streams:
    torch.cuda.current_stream().wait_stream(s)
e1.record()
torch.cuda.synchronize()
print(f"{name} G={G}: {e0.elapsed_time(e1) / (iters // 10):.2f} ms per 10 iterations of all {cpg} chains "
      f"(checksum {sum(float(o.sum()) for o in outs):.12e})", flush=True)
for h in handles:
    lib.rsfm_destroy(h)
