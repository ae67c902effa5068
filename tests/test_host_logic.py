"""Host-side logic on CPU: sharding, pooled-adaptation algebra, R-hat arithmetic, the JSON codec,
facade construction, and the world_size-2 (gloo) reduction path."""
import importlib
import json
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import PKG, ROOT


def test_chain_shard_partition(pkg):
    sh = pkg.ChainShard
    for total, world in ((1024, 1), (1024, 8), (1000, 3), (7, 7), (1048576, 8)):
        parts = [sh.for_rank(total, r, world) for r in range(world)]
        assert parts[0].start == 0 and parts[-1].stop == total
        assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
        counts = [p.count for p in parts]
        assert max(counts) - min(counts) <= 1 and sum(counts) == total
    with pytest.raises(ValueError):
        sh.for_rank(3, 0, 4)
    with pytest.raises(ValueError):
        sh.for_rank(8, 8, 8)


def test_pooled_adaptation_algebra(pkg):
    ad = importlib.import_module(PKG + ".adaptation")
    rng = np.random.default_rng(0)
    for d in (1, 3):
        a = rng.standard_normal((d, d))
        cov = a @ a.T + np.eye(d)
        x = rng.multivariate_normal(np.arange(d) + 1.0, cov, size=5000)
        s = [x.shape[0]] + list(x.sum(axis=0))
        sec = x.T @ x
        s += list(sec[np.tril_indices(d)])
        n, mean, c = ad.moments_from_suffstats(np.array(s), d)
        assert n == 5000 and np.allclose(mean, x.mean(axis=0)) and np.allclose(c, np.cov(x.T), rtol=1e-9)
        fac = ad.proposal_from_suffstats(np.array(s), d)
        target = 2.38 ** 2 / d * np.cov(x.T).reshape(d, d)
        if d == 1:
            assert fac[0] == pytest.approx(target[0, 0], rel=1e-8)        # variance, as the reference stores it
        else:
            low = np.zeros((d, d))
            low[np.tril_indices(d)] = fac
            assert np.allclose(low @ low.T, target, rtol=1e-7)
    # degenerate moments (all samples identical) -> no update, like the swallowed LinAlgError of q4
    assert ad.proposal_from_suffstats(np.array([10.0, 50.0, 250.0]), 1) is None


def test_rhat_from_sums_matches_definition(pkg):
    dg = importlib.import_module(PKG + ".diagnostics")
    rng = np.random.default_rng(1)
    x = rng.standard_normal((40, 300)) + rng.standard_normal((40, 1)) * 0.3       # 40 half-chains
    means, var = x.mean(axis=1), x.var(axis=1, ddof=1)
    n, m = x.shape[1], x.shape[0]
    w = var.mean()
    b_over_n = means.var(ddof=1)
    expect = np.sqrt(((n - 1) / n * w + b_over_n) / w)
    got = dg.rhat_from_sums(m, n, means.sum(), (means ** 2).sum(), var.sum())
    assert got == pytest.approx(expect, rel=1e-12)


def test_json_codec_wire_format(pkg, tmp_path):
    fn = tmp_path / "x.json"
    obj = {"a": np.array([[0.1, 1 / 3, 1350.0]]), "b": [1, 2], "c": np.arange(3)}
    pkg.save_object(obj, str(fn))
    text = fn.read_text()
    # SURVEY 5.4 example bytes
    assert '{"__ndarray__": true, "data": [[0.1, 0.3333333333333333, 1350.0]], "shape": [1, 3]}' in text
    back = pkg.load_object(str(fn))
    assert np.array_equal(back["a"], obj["a"]) and back["a"].dtype == np.float64
    assert back["c"].dtype == np.int64 and back["b"] == [1, 2]
    # errors are printed and swallowed; load returns None (json_save_load.py:177-181)
    assert pkg.load_object(str(tmp_path / "missing.json")) is None
    with pytest.raises(TypeError):
        pkg.numpy_array_encoder(np.float64(1.0))          # NumPy scalars are not handled (reference behaviour)
    assert pkg.numpy_array_decoder({"k": 1}) == {"k": 1}


def test_golden_files_use_reference_wire_format():
    raw = json.load(open(os.path.join(ROOT, "tests", "golden", "sse_grid.json")))
    assert raw["data"]["__ndarray__"] is True and raw["data"]["shape"] == [500]


def test_facade_constructor_mirrors_reference(pkg):
    m = pkg.RateStateModel(number_time_steps=500)
    assert (m.a, m.b, m.mu_ref, m.V_ref, m.k1) == (0.011, 0.014, 0.6, 1.0, 1e-7)
    assert (m.t_start, m.t_final, m.num_tsteps, m.delta_t, m.mu_t_zero) == (0.0, 50.0, 500, 0.1, 0.6)
    assert m.RadiationDamping is True and m.Dc is None
    assert m.num_outputs() == 500
    # quirk q8: float floor can give N - 1
    assert pkg.RateStateModel(number_time_steps=501, end_time=50.1).num_outputs() in (500, 501)
    data = np.zeros(500)
    mc = pkg.MCMC(m, data, 1350.0, ["Uniform", 0.0, 10000.0], 1000.0, nsamples=500)
    assert mc.nburn == 250 and mc.n0 == 0.01 and mc.adapt_interval == 10 and not mc.compat_adapt
    assert mc.qstart_limits.shape == (1, 2)
    assert pkg.MCMC(m, data, 1350.0, {1: 0.0, 2: 1e4}, 1000.0).compat_adapt
    with pytest.raises(ValueError):
        pkg.MCMC(m, data, 1.0, [0, 0, 1], 1.0, param_names=("x",))
    with pytest.raises(TypeError):
        m.evaluate()                                       # Dc not set
    # named parameters of the build: Dc, (a, b, Dc), k1 -- nothing else, and the reference's windowed
    # adaptation exists for one parameter only
    for names in (("Dc",), ("a", "b", "Dc"), ("k1",)):
        assert pkg.MCMC(m, data, 1.0, ["Uniform", 0.0, 1.0], 1.0, param_names=names).param_names == names
    with pytest.raises(ValueError):
        pkg.MCMC(m, data, 1.0, ["Uniform", 0.0, 1.0], 1.0, param_names=("k1", "Dc"))
    with pytest.raises(ValueError):
        pkg.MCMC(m, data, 1.0, {1: 0.0, 2: 1.0}, 1.0, param_names=("a", "b", "Dc"))
    assert pkg.MCMC(m, data, 1.0, {1: 0.0, 2: 0.01}, 1e-3, param_names=("k1",)).compat_adapt


def test_cfg_struct_matches_the_header(pkg):
    """The ctypes mirror of struct rsfm_cfg names every field of include/rsfm.h, in order (a drifted mirror would
    silently shift every later field)."""
    import os
    import re
    from conftest import ROOT
    src = open(os.path.join(ROOT, "include", "rsfm.h")).read()
    body = src[src.index("typedef struct rsfm_cfg {"):src.index("} rsfm_cfg;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.replace("typedef struct rsfm_cfg {", "").strip()
        if not stmt:
            continue
        decl = stmt.split(None, 1)[1] if not stmt.startswith("const") else stmt.split("*", 1)[1]
        for n in decl.split(","):
            names.append(re.sub(r"\[.*\]", "", n).replace("*", "").strip())
    assert names == [f[0] for f in pkg._lib.RsfmCfg._fields_]


def test_qstart_shapes(pkg):
    import torch
    m = pkg.RateStateModel()
    sh = pkg.ChainShard(0, 4, 4)
    mk = lambda q, **kw: pkg.MCMC(m, np.zeros(500), 1.0, [0, 0, 1], q, n_chains=4, **kw)._start_values(
        torch, torch.device("cpu"), kw.get("d", 1) if False else len(kw.get("param_names", ("Dc",))), sh)
    assert mk(5.0).shape == (1, 4)
    assert torch.equal(mk(np.array([1.0, 2.0, 3.0, 4.0]))[0], torch.tensor([1.0, 2.0, 3.0, 4.0], dtype=torch.float64))
    q3 = mk(np.array([0.01, 0.02, 100.0]), param_names=("a", "b", "Dc"))
    assert q3.shape == (3, 4) and q3[2, 3] == 100.0
    with pytest.raises(ValueError):
        mk(np.zeros(5))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_gloo_reductions(tmp_path):
    """The N > 1 host path on CPU: shard ranges per rank, SUM all-reduce of the pooled statistics,
    identical proposal factor on every rank."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import importlib, os, sys, json
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        dist.init_process_group("gloo")
        pkg = importlib.import_module({PKG!r})
        sh = importlib.import_module({PKG!r} + ".sharding")
        ad = importlib.import_module({PKG!r} + ".adaptation")
        rank, world = sh.rank_and_world()
        shard = pkg.ChainShard.for_current_rank(1001)
        rng = np.random.default_rng(7)
        x = rng.normal(1300.0, 60.0, size=(1001, 50))              # all chains, same on both ranks
        mine = x[shard.start:shard.stop]
        s = torch.tensor([mine.size, mine.sum(), (mine ** 2).sum()], dtype=torch.float64)
        sh.all_reduce_sum_(s)
        fac = ad.proposal_from_suffstats(s.numpy(), 1)
        with open(os.path.join({str(tmp_path)!r}, f"rank{{rank}}.json"), "w") as fh:
            json.dump({{"rank": rank, "world": world, "start": shard.start, "stop": shard.stop,
                       "n": s[0].item(), "fac": float(fac[0]), "expect": float(2.38 ** 2 * x.var(ddof=1))}}, fh)
        dist.destroy_process_group()
    """))
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.load(open(tmp_path / f"rank{r}.json")) for r in (0, 1)]
    assert sorted(r["rank"] for r in rows) == [0, 1] and all(r["world"] == 2 for r in rows)
    rows.sort(key=lambda r: r["rank"])
    assert (rows[0]["start"], rows[0]["stop"], rows[1]["start"], rows[1]["stop"]) == (0, 501, 501, 1001)
    assert rows[0]["n"] == rows[1]["n"] == 1001 * 50
    assert rows[0]["fac"] == rows[1]["fac"]
    assert rows[0]["fac"] == pytest.approx(rows[0]["expect"], rel=1e-9)


def test_world_size_2_gloo_pooled_rows_are_sharding_invariant(tmp_path, pkg):
    """The pooled-adaptation exchange on CPU (gloo): group rows (sums over 1,024 chains aligned on the global
    chain id) are all-gathered in rank order and summed sequentially; with shard boundaries on group boundaries
    the pooled moments are the SAME BITS for one and for two ranks (otherwise equal up to summation order)."""
    sh = importlib.import_module(PKG + ".sharding")
    g, rows = pkg._lib.POOL_GROUP, pkg._lib.POOL_ROWS
    assert sh.pool_groups(0, 4096, g) == 4 and sh.pool_groups(2048, 3072, g) == 3 and sh.pool_groups(2148, 3072, g) == 4
    assert sh.max_pool_groups(4096, 2, g) == 2 and sh.max_pool_groups(5000, 2, g) == 3 and sh.max_pool_groups(1 << 20, 8, g) == 128
    script = tmp_path / "w2.py"
    script.write_text(textwrap.dedent(f"""
        import importlib, os, sys, json
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        dist.init_process_group("gloo")
        pkg = importlib.import_module({PKG!r})
        sh = importlib.import_module({PKG!r} + ".sharding")
        G, R = pkg._lib.POOL_GROUP, pkg._lib.POOL_ROWS
        rank, world = sh.rank_and_world()
        out = {{}}
        for total in (4096, 5000):
            rng = np.random.default_rng(total)
            x = rng.normal(1300.0, 60.0, size=total)                    # one draw per chain, same on both ranks
            shard = pkg.ChainShard.for_current_rank(total)
            ng = sh.max_pool_groups(total, world, G)
            loc = torch.zeros((ng, R), dtype=torch.float64)
            first = shard.start // G
            for c in range(shard.start, shard.stop):
                r = c // G - first
                loc[r, 0] += 1.0; loc[r, 1] += x[c]; loc[r, 2] += x[c] * x[c]
            parts = torch.zeros((world * ng, R), dtype=torch.float64)
            sh.all_gather_rows(parts, loc)
            mom = np.zeros(3)
            for p in parts.numpy():
                mom += p[:3]
            out[str(total)] = [float(v).hex() for v in mom]
        json.dump(out, open(os.path.join({str(tmp_path)!r}, f"rows_rank{{rank}}.json"), "w"))
        dist.destroy_process_group()
    """))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert run.returncode == 0, run.stderr[-2000:]
    got = [json.load(open(tmp_path / f"rows_rank{r}.json")) for r in (0, 1)]
    assert got[0] == got[1]                                              # every rank holds the same moments
    for total in (4096, 5000):
        x = np.random.default_rng(total).normal(1300.0, 60.0, size=total)
        mom = np.zeros(3)
        for g0 in range(0, total, g):                                    # world = 1: the same group rows, in order
            row = np.zeros(3)
            for c in range(g0, min(g0 + g, total)):
                row += np.array([1.0, x[c], x[c] * x[c]])
            mom += row
        one = [float(v).hex() for v in mom]
        if total % (2 * g) == 0:
            assert got[0][str(total)] == one                             # aligned shards: bit-identical
        else:
            two = np.array([float.fromhex(v) for v in got[0][str(total)]])
            assert two[0] == total and np.allclose(two, mom, rtol=1e-13)
