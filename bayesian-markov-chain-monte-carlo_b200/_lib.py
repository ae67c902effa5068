"""ctypes binding of librsfm.so (the C ABI declared in include/rsfm.h).

There is no CPU fallback anywhere in this package: if the CUDA library has not
been built, or no B200 is visible, every compute entry point raises.
Build with ``python -c "import __graft_entry__ as g; g.build()"`` at the repo root.
"""
import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "librsfm.so")

RSFM_MAX_PARAMS = 3

LOAD_SINE_DECAY, LOAD_VSTEP, LOAD_TABLE = 0, 1, 2
LAW_AGING, LAW_SLIP = 0, 1
INTEG_PARITY, INTEG_CARRY = 0, 1
ADAPT_NONE, ADAPT_COMPAT, ADAPT_POOLED = 0, 1, 2
CHAIN_OK, CHAIN_NMAX, CHAIN_HSMALL, CHAIN_NONFINITE = 0, 2, 3, 4
OBS_ACC, OBS_MU = 0, 1
VARIANT_AUTO, VARIANT_DEFAULT, VARIANT_STIFF = 0, 1, 2
POOL_GROUP, POOL_ROWS = 1024, 16
PARAM_DC, PARAM_K1 = 0, 1
ABI_VERSION = 4

# every symbol include/rsfm.h declares (tests check the library exports them all)
EXPORTED_SYMBOLS = (
    "rsfm_abi_version", "rsfm_last_error", "rsfm_cfg_defaults", "rsfm_device_count",
    "rsfm_forward_batch", "rsfm_create", "rsfm_destroy", "rsfm_init", "rsfm_run",
    "rsfm_run_deterministic", "rsfm_spec_depth", "rsfm_get_state", "rsfm_set_state", "rsfm_iteration",
    "rsfm_get_totals", "rsfm_get_suffstats", "rsfm_set_proposal_chol", "rsfm_chain_diagnostics",
    "rsfm_kde_grid", "rsfm_measure_fp64_peak", "rsfm_trim",
    "rsfm_pooled_groups", "rsfm_pooled_partials", "rsfm_pooled_update",
    "rsfm_philox_raw", "rsfm_philox_draws", "rsfm_rhs_eval", "rsfm_get_ring", "rsfm_set_ring",
    "rsfm_chain_groups", "rsfm_join", "rsfm_pooled_install",
)


class RsfmCfg(C.Structure):
    """Mirror of ``struct rsfm_cfg`` (include/rsfm.h)."""
    _fields_ = [
        ("a", C.c_double), ("b", C.c_double), ("mu_ref", C.c_double), ("V_ref", C.c_double),
        ("k1", C.c_double),
        ("t_start", C.c_double), ("t_final", C.c_double), ("delta_t", C.c_double),
        ("mu_t_zero", C.c_double),
        ("vstep_period", C.c_double), ("vstep_factor", C.c_double),
        ("rtol", C.c_double), ("atol", C.c_double), ("n0", C.c_double),
        ("lo", C.c_double * RSFM_MAX_PARAMS), ("hi", C.c_double * RSFM_MAX_PARAMS),
        ("n_out", C.c_int32), ("nmax", C.c_int32), ("radiation_damping", C.c_int32),
        ("loading", C.c_int32), ("integ_mode", C.c_int32), ("n_params", C.c_int32),
        ("n_prior_len", C.c_int32), ("adapt_interval", C.c_int32), ("adapt_mode", C.c_int32),
        ("spec_depth", C.c_int32),
        ("observable", C.c_int32), ("solver_variant", C.c_int32), ("stiff_exact", C.c_int32),
        ("block_threads", C.c_int32), ("chain_groups", C.c_int32), ("round_packing", C.c_int32),
        ("state_law", C.c_int32), ("n_load_table", C.c_int32), ("load_dt", C.c_double), ("load_table_dev", C.c_void_p),
        ("sampled_param", C.c_int32), ("dc_fixed", C.c_double),
    ]


class RsfmError(RuntimeError):
    pass


_lib = None


def load():
    """Load librsfm.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RsfmError(
            f"{LIB_PATH} not found: the CUDA extension is not built and this package has no CPU "
            "fallback.  Run `python -c \"import __graft_entry__ as g; g.build()\"` at the repo root.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    cfgp = C.POINTER(RsfmCfg)
    lib.rsfm_abi_version.restype = C.c_int
    lib.rsfm_last_error.restype = C.c_char_p
    lib.rsfm_cfg_defaults.argtypes = [cfgp]
    lib.rsfm_cfg_defaults.restype = None
    lib.rsfm_device_count.restype = C.c_int
    lib.rsfm_forward_batch.argtypes = [cfgp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.rsfm_forward_batch.restype = C.c_int
    lib.rsfm_create.argtypes = [cfgp, i32, u64, u64]
    lib.rsfm_create.restype = vp
    lib.rsfm_destroy.argtypes = [vp]
    lib.rsfm_destroy.restype = None
    lib.rsfm_init.argtypes = [vp, vp, vp, vp]
    lib.rsfm_init.restype = C.c_int
    lib.rsfm_run.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.rsfm_run.restype = C.c_int
    lib.rsfm_spec_depth.argtypes = [vp]
    lib.rsfm_spec_depth.restype = C.c_int
    lib.rsfm_chain_groups.argtypes = [vp]
    lib.rsfm_chain_groups.restype = C.c_int
    lib.rsfm_join.argtypes = [vp, vp]
    lib.rsfm_join.restype = C.c_int
    lib.rsfm_pooled_install.argtypes = [vp, vp]
    lib.rsfm_pooled_install.restype = C.c_int
    lib.rsfm_run_deterministic.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.rsfm_run_deterministic.restype = C.c_int
    lib.rsfm_get_state.argtypes = [vp] + [vp] * 8 + [vp]
    lib.rsfm_get_state.restype = C.c_int
    lib.rsfm_set_state.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    lib.rsfm_set_state.restype = C.c_int
    lib.rsfm_iteration.argtypes = [vp]
    lib.rsfm_iteration.restype = i64
    lib.rsfm_get_totals.argtypes = [vp, C.POINTER(u64), vp]
    lib.rsfm_get_totals.restype = C.c_int
    lib.rsfm_get_suffstats.argtypes = [vp, vp, i32, vp]
    lib.rsfm_get_suffstats.restype = C.c_int
    lib.rsfm_set_proposal_chol.argtypes = [vp, C.POINTER(dbl), vp]
    lib.rsfm_set_proposal_chol.restype = C.c_int
    lib.rsfm_chain_diagnostics.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.rsfm_chain_diagnostics.restype = C.c_int
    lib.rsfm_kde_grid.argtypes = [vp, i64, vp, i32, dbl, vp, vp]
    lib.rsfm_kde_grid.restype = C.c_int
    lib.rsfm_measure_fp64_peak.argtypes = [dbl, C.POINTER(dbl)]
    lib.rsfm_measure_fp64_peak.restype = C.c_int
    lib.rsfm_trim.argtypes = []
    lib.rsfm_trim.restype = C.c_int
    lib.rsfm_get_ring.argtypes = [vp, vp, vp]
    lib.rsfm_get_ring.restype = C.c_int
    lib.rsfm_set_ring.argtypes = [vp, vp, vp]
    lib.rsfm_set_ring.restype = C.c_int
    lib.rsfm_pooled_groups.argtypes = [vp]
    lib.rsfm_pooled_groups.restype = C.c_int
    lib.rsfm_pooled_partials.argtypes = [vp, vp, i32, vp]
    lib.rsfm_pooled_partials.restype = C.c_int
    lib.rsfm_pooled_update.argtypes = [vp, vp, i32, vp, i32, i32, vp, vp]
    lib.rsfm_pooled_update.restype = C.c_int
    lib.rsfm_philox_raw.argtypes = [vp, vp, i32, vp]
    lib.rsfm_philox_raw.restype = C.c_int
    lib.rsfm_philox_draws.argtypes = [u64, u64, i32, C.c_uint32, i32, dbl, vp, vp]
    lib.rsfm_philox_draws.restype = C.c_int
    lib.rsfm_rhs_eval.argtypes = [cfgp, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp]
    lib.rsfm_rhs_eval.restype = C.c_int
    if lib.rsfm_abi_version() != ABI_VERSION:
        raise RsfmError(f"librsfm ABI version {lib.rsfm_abi_version()} != {ABI_VERSION}: rebuild the library")
    _lib = lib
    return lib


def check(rc: int, what: str = "librsfm") -> None:
    if rc != 0:
        msg = load().rsfm_last_error().decode("utf-8", "replace")
        raise RsfmError(f"{what} failed (code {rc}): {msg}")


def default_cfg() -> RsfmCfg:
    cfg = RsfmCfg()
    load().rsfm_cfg_defaults(C.byref(cfg))
    return cfg


def require_cuda():
    """Import torch, insist on a visible CUDA device, and return the torch module."""
    import torch
    if not torch.cuda.is_available():
        raise RsfmError("no CUDA device visible: the RSF-MCMC path is CUDA-only (sm_100a), "
                        "there is no CPU fallback")
    load()
    return torch


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream(torch, device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
