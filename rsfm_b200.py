"""Importable alias of the package directory ``bayesian-markov-chain-monte-carlo_b200``
(hyphens are not valid in an ``import`` statement).  ``import rsfm_b200`` returns
that package itself."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
