"""Import shim for the UNMODIFIED reference (test infrastructure, container-only).

The reference's ``imports.py:10-12`` imports matplotlib and mysql.connector
unconditionally; neither is installed here.  Registering empty stub modules
before importing lets ``RateStateModel.evaluate`` (RateStateModel.py:188) and
``MCMC.sample(False)`` (MCMC.py:391) run unmodified.

``/root/reference`` exists only in the build container, never on the GPU box:
nothing under tests/ -m gpu, smoke() or bench.py may import this module.  It is
used solely by ``oracle/make_golden.py`` to generate ``tests/golden/*.json`` and
by CPU-side tests that skip when the reference is absent.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("RSFM_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "MCMC.py"))


def install_stubs() -> None:
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation",
                 "mysql", "mysql.connector"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.animation"].FuncAnimation = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["mysql"].connector = sys.modules["mysql.connector"]


def load_reference():
    """Return (RateStateModel_module, MCMC_module) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    install_stubs()
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import RateStateModel as ref_rsm  # noqa: E402
    import MCMC as ref_mcmc  # noqa: E402
    return ref_rsm, ref_mcmc
