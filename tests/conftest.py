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
