"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT the product).

Wraps ``oracle/librsf_oracle.so`` (built by ``make -C oracle`` from
rsf_oracle.c, a plain-C restatement of RateStateModel.py:188-395, MCMC.py:129-544
and SciPy's dop853).  Importers allowed: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference legs.  Nothing in the product package
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librsf_oracle.so")

LOAD_SINE_DECAY = 0
LOAD_VSTEP = 1
LOAD_TABLE = 2
OBS_ACC = 0
OBS_MU = 1
LAW_AGING = 0
LAW_SLIP = 1
PARAM_DC = 0
PARAM_K1 = 1


class OrcModel(C.Structure):
    _fields_ = [
        ("a", C.c_double), ("b", C.c_double), ("mu_ref", C.c_double),
        ("V_ref", C.c_double), ("k1", C.c_double), ("Dc", C.c_double),
        ("t_start", C.c_double), ("t_final", C.c_double),
        ("num_tsteps", C.c_int),
        ("mu_t_zero", C.c_double),
        ("radiation_damping", C.c_int),
        ("loading", C.c_int),
        ("vstep_period", C.c_double), ("vstep_factor", C.c_double),
        ("rtol", C.c_double), ("atol", C.c_double),
        ("nmax", C.c_int),
        ("observable", C.c_int),
        ("state_law", C.c_int),
        ("load_table", C.POINTER(C.c_double)), ("n_load_table", C.c_int), ("load_dt", C.c_double),
        ("sampled_param", C.c_int),
    ]


class OrcStats(C.Structure):
    _fields_ = [
        ("nrhs", C.c_int64), ("nstep", C.c_int64), ("naccpt", C.c_int64),
        ("nrejct", C.c_int64), ("filled", C.c_int), ("istate", C.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle library if needed and return its path."""
    src = os.path.join(_HERE, "rsf_oracle.c")
    import shutil
    stale = not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if (force or stale) and shutil.which("make") and shutil.which("gcc"):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    if not os.path.exists(_SO):
        raise RuntimeError("oracle library missing and no compiler available to build it")
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        _lib.orc_model_defaults.argtypes = [C.POINTER(OrcModel)]
        _lib.orc_rhs.argtypes = [C.POINTER(OrcModel), C.c_double, dp, dp]
        _lib.orc_forward.argtypes = [C.POINTER(OrcModel), dp, dp, dp, dp, dp, C.POINTER(OrcStats)]
        _lib.orc_forward.restype = C.c_int
        _lib.orc_sse.argtypes = [dp, dp, C.c_int]
        _lib.orc_sse.restype = C.c_double
        _lib.orc_forward_batch.argtypes = [C.POINTER(OrcModel), dp, C.c_int, dp, C.c_int, dp, dp,
                                           C.POINTER(C.c_int64), C.c_int]
        _lib.orc_forward_batch.restype = C.c_int
        _lib.orc_chain_replay.argtypes = [C.POINTER(OrcModel), dp, C.c_int, C.c_double, C.c_double,
                                          C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp,
                                          dp, dp, C.POINTER(C.c_uint8), dp, C.POINTER(C.c_int64)]
        _lib.orc_chain_replay.restype = C.c_int
        _lib.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                           C.POINTER(C.c_uint32)]
        _lib.orc_chain_replay_nd.argtypes = [C.POINTER(OrcModel), dp, C.c_int, C.c_int, dp, dp, dp, C.c_int, C.c_int,
                                             dp, dp, dp, dp, dp, C.POINTER(C.c_uint8), C.POINTER(C.c_int64)]
        _lib.orc_chain_replay_nd.restype = C.c_int
        _lib.orc_philox_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, dp]
        _lib.orc_philox_draws.restype = None
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def make_model(Dc=1000.0, number_time_steps=500, start_time=0.0, end_time=50.0, **kw) -> OrcModel:
    m = OrcModel()
    lib().orc_model_defaults(C.byref(m))
    m.Dc = float(Dc)
    m.num_tsteps = int(number_time_steps)
    m.t_start = float(start_time)
    m.t_final = float(end_time)
    for k, v in kw.items():
        if not hasattr(m, k):
            raise AttributeError(k)
        if k == "load_table":
            tab = np.ascontiguousarray(v, dtype=np.float64)
            m._load_table_keepalive = tab                     # the struct only holds the pointer
            m.load_table = tab.ctypes.data_as(C.POINTER(C.c_double))
            m.n_load_table = tab.size
            continue
        setattr(m, k, v)
    return m


def num_outputs(m: OrcModel) -> int:
    """int(floor((t_final - t_start)/delta_t)), RateStateModel.py:358."""
    delta_t = (m.t_final - m.t_start) / m.num_tsteps
    return int(np.floor((m.t_final - m.t_start) / delta_t))


def rhs(m: OrcModel, t: float, y) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros(3)
    lib().orc_rhs(C.byref(m), float(t), _dp(y), _dp(out))
    return out


def forward(m: OrcModel, full: bool = False):
    """RateStateModel.evaluate() without the noise draw.

    Returns (t, acc, stats) or, with full=True, (t, mu, theta, vel, acc, stats).
    """
    n = num_outputs(m)
    t, mu, th, vel, acc = (np.zeros(n) for _ in range(5))
    st = OrcStats()
    got = lib().orc_forward(C.byref(m), _dp(t), _dp(mu), _dp(th), _dp(vel), _dp(acc), C.byref(st))
    assert got == n
    if full:
        return t, mu, th, vel, acc, st
    return t, acc, st


def sse(acc, data) -> float:
    acc = np.ascontiguousarray(acc, dtype=np.float64)
    data = np.ascontiguousarray(data, dtype=np.float64)
    return lib().orc_sse(_dp(acc), _dp(data), acc.size)


def forward_batch(m: OrcModel, dc, data=None, want_acc=False, nthreads=1):
    dc = np.ascontiguousarray(dc, dtype=np.float64)
    n = num_outputs(m)
    sse_out = np.zeros(dc.size) if data is not None else None
    acc_out = np.zeros((dc.size, n)) if want_acc else None
    nrhs = np.zeros(dc.size, dtype=np.int64)
    if data is not None:
        data = np.ascontiguousarray(data, dtype=np.float64)
        assert data.size == n
    rc = lib().orc_forward_batch(C.byref(m), _dp(dc), dc.size, _dp(data), n, _dp(sse_out), _dp(acc_out),
                                 nrhs.ctypes.data_as(C.POINTER(C.c_int64)), int(nthreads))
    if rc != 0:
        raise RuntimeError("oracle forward_batch failed")
    return sse_out, acc_out, nrhs


def chain_replay(m: OrcModel, data, qstart, lo, hi, n_prior_len, nsamples, proposals, uniforms,
                 gammas, adapt_interval=10, compat_adapt=False):
    data = np.ascontiguousarray(data, dtype=np.float64)
    proposals = np.ascontiguousarray(proposals, dtype=np.float64)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    gammas = np.ascontiguousarray(gammas, dtype=np.float64)
    chain = np.zeros(nsamples + 1)
    s2 = np.zeros(nsamples + 1)
    acc = np.zeros(nsamples, dtype=np.uint8)
    vstart = C.c_double()
    nsolves = C.c_int64()
    rc = lib().orc_chain_replay(C.byref(m), _dp(data), data.size, float(qstart), float(lo), float(hi),
                                int(n_prior_len), int(nsamples), int(adapt_interval),
                                int(bool(compat_adapt)), _dp(proposals), _dp(uniforms), _dp(gammas),
                                _dp(chain), _dp(s2), acc.ctypes.data_as(C.POINTER(C.c_uint8)),
                                C.byref(vstart), C.byref(nsolves))
    if rc != 0:
        raise RuntimeError("oracle chain_replay failed")
    return chain, s2, acc, vstart.value, nsolves.value


def philox4x32_10(ctr, key) -> np.ndarray:
    c = (C.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return np.array(list(o), dtype=np.uint32)


def chain_replay_nd(m: OrcModel, data, qstart, lo, hi, n_prior_len, nsamples, proposals, uniforms, gammas):
    """d-parameter replay (d = 1: Dc; d = 3: a, b, Dc) with ABSOLUTE proposals [nsamples, d] and per-parameter
    bounds.  Returns (chain [nsamples+1, d], s2 [nsamples+1], accepts [nsamples], nsolves)."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    qstart = np.ascontiguousarray(np.atleast_1d(qstart), dtype=np.float64)
    d = qstart.size
    lo = np.ascontiguousarray(np.broadcast_to(lo, (d,)), dtype=np.float64)
    hi = np.ascontiguousarray(np.broadcast_to(hi, (d,)), dtype=np.float64)
    proposals = np.ascontiguousarray(proposals, dtype=np.float64).reshape(nsamples, d)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    gammas = np.ascontiguousarray(gammas, dtype=np.float64)
    chain = np.zeros((nsamples + 1, d))
    s2 = np.zeros(nsamples + 1)
    acc = np.zeros(nsamples, dtype=np.uint8)
    nsolves = C.c_int64()
    rc = lib().orc_chain_replay_nd(C.byref(m), _dp(data), data.size, d, _dp(qstart), _dp(lo), _dp(hi),
                                   int(n_prior_len), int(nsamples), _dp(proposals), _dp(uniforms), _dp(gammas),
                                   _dp(chain), _dp(s2), acc.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nsolves))
    if rc != 0:
        raise RuntimeError(f"oracle chain_replay_nd failed ({rc})")
    return chain, s2, acc, nsolves.value


def philox_draws(seed: int, chain: int, iteration: int, shape: float) -> np.ndarray:
    """(z0, z1, z2, U, unit gamma, gamma attempts) of one chain and iteration on the CUDA path's Philox stream."""
    out = np.zeros(6)
    lib().orc_philox_draws(C.c_uint64(seed), C.c_uint64(chain), C.c_uint32(iteration), float(shape), _dp(out))
    return out
