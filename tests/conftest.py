import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "bayesian-markov-chain-monte-carlo_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _decode(d):
    if d.get("__ndarray__"):
        return np.array(d["data"]).reshape(d["shape"])
    return d


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f, object_hook=_decode)


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def built_lib():
    """librsfm.so, built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return importlib.import_module(PKG)._lib.load()


@pytest.fixture(scope="session")
def cuda(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def oracle_ulp_floor(orc, dc, **model_kw):
    """Reproducibility floor of the reference solver itself: the largest change of the CPU oracle's trajectory
    under a one-to-four-ulp perturbation of Dc, relative to max|acc|.  Outside the stiff regime this is ~1e-13
    (rounding); in the stability-limited regime (Dc < ~1, velocity steps) the accepted-step sequence changes and
    the floor sits at the solver's own tolerance, 2-4e-6 of the trajectory scale (rtol = 1e-6 on V, divided by
    delta_t).  Any two correct implementations of the same DOP853 -- SciPy and the C oracle included, see
    tests/test_oracle_vs_scipy.py -- differ by about this much there, so parity gates in that regime are stated
    as a multiple of this measured floor."""
    base = orc.forward(orc.make_model(Dc=dc, **model_kw))[1]
    scale = float(np.max(np.abs(base)))
    floor = 0.0
    for eps in (1.2e-16, -1.2e-16, 2.3e-16, -2.3e-16, 4.5e-16, -4.5e-16):
        pert = orc.forward(orc.make_model(Dc=dc * (1.0 + eps), **model_kw))[1]
        floor = max(floor, float(np.max(np.abs(pert - base))) / scale)
    return base, scale, floor


def rhs_term_scales(row, a=0.011, b=0.014, mu_ref=0.6, v_ref=1.0):
    """Magnitude of the terms each component of friction(t, y) is a difference of (RateStateModel.py:336-353):
    theta' = 1 - v theta/Dc, mu' = k'V_l - k'v, V' = (v/a)(mu' - (b/theta) theta').  Two correct evaluations
    agree to a few ulp OF THESE TERMS; near sliding steady state the differences themselves are 1e-4..1e-6 of
    them, so 'relative to the result' would measure the cancellation, not the arithmetic.  row = (damping, Dc,
    t, mu, theta, V, ...)."""
    dc, t, mu, th = row[1], row[2], row[3], row[4]
    kp = 1e-2 * 10 / dc
    v_l = v_ref * (1 + np.exp(-t / 20) * np.sin(10 * t))
    temp = (mu - mu_ref - b * np.log(v_ref * th / dc)) / a
    v = v_ref * np.exp(temp)
    # the exponent carries (|mu - mu_ref| + b |log|)/a of rounding amplification into v
    amp = 1.0 + (abs(mu - mu_ref) + b * abs(np.log(v_ref * th / dc))) / a
    s_th = (1.0 + v * th / dc) * amp
    s_mu = kp * (abs(v_l) + v) * amp
    s_v = v / a * (s_mu + b / th * s_th)
    return np.array([s_mu, s_th, s_v])
