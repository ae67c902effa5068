"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/rsfm.h declares.
No compute calls: this runs on the CPU-only build container."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "rsfm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsfm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(built_lib, pkg):
    declared = _declared_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(built_lib, name), f"librsfm.so does not export {name}"
    assert sorted(pkg._lib.EXPORTED_SYMBOLS) == declared


def test_abi_version_and_defaults(built_lib, pkg):
    assert built_lib.rsfm_abi_version() == 4
    cfg = pkg._lib.default_cfg()
    # RateStateModel.py:5-11 / :374 / MCMC.py:97
    assert (cfg.a, cfg.b, cfg.mu_ref, cfg.V_ref, cfg.k1) == (0.011, 0.014, 0.6, 1.0, 1e-7)
    assert (cfg.t_start, cfg.t_final, cfg.n_out, cfg.delta_t) == (0.0, 50.0, 500, 0.1)
    assert (cfg.rtol, cfg.atol, cfg.nmax, cfg.n0) == (1e-6, 1e-10, 500, 0.01)
    assert cfg.radiation_damping == 1 and cfg.loading == 0 and cfg.integ_mode == 0
    assert cfg.spec_depth == 0 and cfg.adapt_interval == 10
    assert cfg.sampled_param == 0 and cfg.dc_fixed == 0.0            # the chain samples Dc (MCMC.py:381)


def test_sass_is_sm100a_with_tma_bulk_copy(built_lib, pkg):
    """The shipped cubin targets sm_100a and stages the series with the TMA bulk-copy engine."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out            # cp.async.bulk
    assert "DFMA" in out


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = pkg.RateStateModel()
    m.Dc = 1000.0
    with pytest.raises(pkg._lib.RsfmError):
        m.evaluate()
    with pytest.raises(pkg._lib.RsfmError):
        pkg.MCMC(m, [0.0] * 500, 1000.0, ["Uniform", 0.0, 1e4], 1000.0, nsamples=4).sample(False)


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "bayesian-markov-chain-monte-carlo_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in text and "import oracle" not in text, f
                assert "rsf_oracle" not in text, f


def test_cfg_validation_needs_no_device(built_lib, pkg):
    """Argument checks come before any device work: bad configurations are refused with a message,
    with or without a GPU (rsfm_status RSFM_ERR_INVALID, like the ValueError a Python reference raises)."""
    import ctypes as C
    lib = pkg._lib.load()
    cfg = pkg._lib.RsfmCfg()
    lib.rsfm_cfg_defaults(C.byref(cfg))

    def refused(**kw):
        c = pkg._lib.RsfmCfg()
        C.memmove(C.byref(c), C.byref(cfg), C.sizeof(c))
        for k, v in kw.items():
            setattr(c, k, v)
        rc = lib.rsfm_forward_batch(C.byref(c), 4, None, None, None, None, None, None, None, None, None, None, None, None)
        return rc, lib.rsfm_last_error().decode()

    for kw, word in (({"loading": pkg._lib.LOAD_VSTEP, "vstep_period": 0.0}, "vstep_period"),
                     ({"loading": pkg._lib.LOAD_VSTEP, "vstep_factor": 0.0}, "vstep_factor"),
                     ({"rtol": 0.0}, "rtol"), ({"n_params": 2}, "n_params"), ({"loading": 7}, "loading"),
                     ({"adapt_interval": 1, "adapt_mode": pkg._lib.ADAPT_COMPAT}, "adapt_interval"),
                     ({"observable": 5}, "observable"), ({"solver_variant": 9}, "solver_variant"),
                     ({"state_law": 3}, "state_law"), ({"round_packing": 2}, "round_packing"), ({"chain_groups": 9}, "chain_groups"), ({"loading": 2}, "RSFM_LOAD_TABLE"),
                     ({"block_threads": 48}, "block_threads"), ({"spec_depth": 6}, "spec_depth"),
                     ({"delta_t": 0.0}, "grid"), ({"a": -1.0}, "positive")):
        rc, msg = refused(**kw)
        assert rc < 0 and word in msg, (kw, rc, msg)
    # the reference accepts any adapt_interval and never uses it with list priors (MCMC.py:58, q2)
    for kw in ({"adapt_interval": 1}, {"adapt_interval": 100, "adapt_mode": pkg._lib.ADAPT_POOLED}):
        rc, msg = refused(**kw)
        assert rc < 0 and "dc_dev" in msg, (kw, msg)
    rc, msg = refused()                        # valid cfg, NULL parameter pointer
    assert rc < 0 and "dc_dev" in msg
