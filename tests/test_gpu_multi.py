"""Multi-GPU identity on hardware ("multi-GPU" row): runs of 1 and 2 ranks (NCCL) give identical
per-chain streams AND identical pooled statistics -- hence identical adaptive chains -- because the Philox counter
carries the global chain id and the pooled moments are summed from group rows in global chain order.
Skipped on a single-GPU box (`gpurun --gpus 2` runs it; the log is kept in profiles/)."""
import json
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import PKG, ROOT, load_golden

pytestmark = pytest.mark.gpu

WORKER = """
import importlib, os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.barrier(); os.dup2(saved, 1)
pkg = importlib.import_module({pkg!r})
cases = json.load(open(os.path.join({tmp!r}, "cases.json")))
data = np.load(os.path.join({tmp!r}, "data.npy"))
for name, kw in cases.items():
    q0 = np.load(os.path.join({tmp!r}, name + "_q0.npy"))
    kw = dict(kw); kw["param_names"] = tuple(kw["param_names"])
    mc = pkg.MCMC(pkg.RateStateModel(), data, 1350.0, ["Uniform", 0.0, 1e4], q0, verbose=False, shard=True, **kw)
    out = mc.sample(False)
    diag = mc.diagnostics()
    np.save(os.path.join({tmp!r}, f"{{name}}_rank{{rank}}.npy"), np.ascontiguousarray(out))
    hist = [[int(e)] + [float(x) for x in f] for e, f in getattr(mc, "adapt_history", [])]
    json.dump({{"hist": hist, "rhat": diag["rhat"], "ess": diag["ess"], "mean": diag["mean"], "stats": {{k: v for k, v in mc.stats.items() if not isinstance(v, float)}},
               "start": mc.stats["chain_id0"], "count": mc.stats["n_chains_local"]}},
              open(os.path.join({tmp!r}, f"{{name}}_rank{{rank}}.json"), "w"))
dist.barrier()
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_one_and_two_ranks_give_identical_chains_and_pooled_statistics(cuda, pkg, tmp_path, world):
    torch = cuda
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    g = load_golden("sse_grid.json")
    np.save(tmp_path / "data.npy", g["data"])
    rng = np.random.default_rng(0)
    # 16,384 per rank: the pooled d = 3 case runs with chain groups (2 per rank, 4 on the single GPU); the d = 1 cases run
    # the speculative kernel on the ranks (two lanes per chain, no groups) and one thread per chain (4 groups) on the
    # single GPU -- the chains must still be the same, bit for bit
    c = 32768
    cases = {
        "dc_fixed": dict(nsamples=24, n_chains=c, seed=5, param_names=["Dc"]),
        "dc_pooled": dict(nsamples=60, n_chains=c, seed=5, param_names=["Dc"], adapt="pooled", adapt_start=20),
        "abdc_pooled": dict(nsamples=60, n_chains=c, seed=6, param_names=["a", "b", "Dc"], adapt="pooled", adapt_start=20,
                            bounds=[[0.0100, 0.0120], [0.0130, 0.0150], [800.0, 2200.0]]),
    }
    q0 = {"dc_fixed": rng.uniform(900.0, 2000.0, c), "dc_pooled": rng.uniform(900.0, 2000.0, c),
          "abdc_pooled": np.stack([rng.uniform(0.0105, 0.0115, c), rng.uniform(0.0135, 0.0145, c), rng.uniform(1000.0, 1800.0, c)], axis=1)}
    for k, v in q0.items():
        np.save(tmp_path / f"{k}_q0.npy", v)
    json.dump(cases, open(tmp_path / "cases.json", "w"))
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, pkg=PKG, tmp=str(tmp_path)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr[-3000:]
    for name, kw in cases.items():
        kw = dict(kw)
        kw["param_names"] = tuple(kw["param_names"])
        mc = pkg.MCMC(pkg.RateStateModel(), g["data"], 1350.0, ["Uniform", 0.0, 1e4], q0[name], verbose=False, **kw)
        full = mc.sample(False)
        diag = mc.diagnostics()
        parts = [np.load(tmp_path / f"{name}_rank{r}.npy") for r in range(world)]
        meta = [json.load(open(tmp_path / f"{name}_rank{r}.json")) for r in range(world)]
        assert [m["start"] for m in meta] == [r * c // world for r in range(world)]
        assert np.array_equal(np.concatenate(parts, axis=0), full), name            # identical chains, bit for bit
        if "adapt" in kw:
            hist = [[int(e)] + [float(x) for x in f] for e, f in mc.adapt_history]
            assert len(hist) >= 3 and all(m["hist"] == hist for m in meta), name    # identical pooled factors
            assert all(m["stats"]["pool_rows_gathered"] == world * (c // world // 1024) for m in meta)
            per_rank = 2 if len(kw["param_names"]) == 3 else 1
            assert all(m["stats"]["chain_groups"] == per_rank for m in meta) and mc.stats["chain_groups"] == 4
        for m in meta:                                                              # R-hat / ESS sums all-reduced
            assert np.allclose(m["rhat"], diag["rhat"], rtol=1e-10) and np.allclose(m["ess"], diag["ess"], rtol=1e-10)
            assert np.allclose(m["mean"], diag["mean"], rtol=1e-12)
