"""B200-native RSF-MCMC hot path: drop-in ``RateStateModel`` / ``MCMC`` classes over librsfm.so.

The directory name contains hyphens (it is fixed by the project layout), so
import it with ``importlib.import_module("bayesian-markov-chain-monte-carlo_b200")`` or
through the alias module ``rsfm_b200`` at the repo root.
"""
from .rate_state_model import RateStateModel, A, B, MU_REF, V_REF, K1, START_TIME, END_TIME
from .sampler import MCMC
from .ndarray_json import save_object, load_object, numpy_array_encoder, numpy_array_decoder
from .sharding import ChainShard
from .driver import RSF, measure_execution_time
from .posterior import gaussian_kde_pdf
from . import _lib

__all__ = ["RateStateModel", "MCMC", "RSF", "measure_execution_time", "gaussian_kde_pdf", "save_object", "load_object", "numpy_array_encoder",
           "numpy_array_decoder", "ChainShard", "A", "B", "MU_REF", "V_REF", "K1", "START_TIME", "END_TIME"]
