"""``RateStateModel`` -- drop-in for the reference forward model, computed on B200.

Mirrors the reference class (RateStateModel.py:15-395): same constructor, same
public mutable attributes, ``evaluate() -> (t, acc, acc_noise)`` with float64
arrays of length ``num_steps``.  The ODE solve itself runs in librsfm's
``rsf_forward_kernel`` (one CUDA thread per parameter set); nothing here
integrates on the CPU.

Extensions (absent from the reference): ``evaluate_batch`` for many parameter
sets at once, ``loading`` / ``integ_mode`` selectors, and
``observable``: ``"acc"`` (reference: the backward-difference
acceleration, :388) or ``"mu"`` (the friction series the reference integrates and
stores, :385, but never compares with data) -- what ``evaluate()[1]``, ``evaluate_batch``
and the sampler's sum of squares are computed on.
"""
import ctypes as C

import numpy as np

from . import _lib

# reference module constants, RateStateModel.py:5-11
A = 0.011
B = 0.014
MU_REF = 0.6
V_REF = 1.0
K1 = 1.0e-7
START_TIME = 0.0
END_TIME = 50.0

_LOADING = {"sine_decay": _lib.LOAD_SINE_DECAY, "vstep": _lib.LOAD_VSTEP, "table": _lib.LOAD_TABLE}
_LAW = {"aging": _lib.LAW_AGING, "slip": _lib.LAW_SLIP}
_INTEG = {"parity": _lib.INTEG_PARITY, "carry": _lib.INTEG_CARRY}
_OBSERVABLE = {"acc": _lib.OBS_ACC, "mu": _lib.OBS_MU}
_VARIANT = {"auto": _lib.VARIANT_AUTO, "default": _lib.VARIANT_DEFAULT, "stiff": _lib.VARIANT_STIFF}


class RateStateModel:
    """Rate-and-state friction spring-slider (Dieterich aging law).

    Reference: RateStateModel.py:109-186 (constructor / attributes) and
    :188-395 (``evaluate``).
    """

    def __init__(self, number_time_steps=500, start_time=START_TIME, end_time=END_TIME):
        # RateStateModel.py:167-184
        self.a = A
        self.b = B
        self.mu_ref = MU_REF
        self.V_ref = V_REF
        self.k1 = K1
        self.t_start = start_time
        self.t_final = end_time
        self.num_tsteps = number_time_steps
        self.delta_t = (end_time - start_time) / number_time_steps
        self.mu_t_zero = MU_REF
        self.RadiationDamping = True
        self.Dc = None
        # extensions
        self.loading = "sine_decay"
        self.vstep_period = 1000.0
        self.vstep_factor = 10.0
        self.integ_mode = "parity"
        self.observable = "acc"
        self.state_law = "aging"       # "slip": Ruina's law instead of Dieterich's (RateStateModel.py:340); extension
        self.load_table = None         # loading = "table": V_l/V_ref - 1 at t_start + i * load_dt (piecewise linear)
        self.load_dt = None            #   spacing of the table; None = delta_t
        self._load_table_dev = None
        self.solver_variant = "auto"   # "default" / "stiff": force a kernel variant (tests, tuning)
        self.stiff_exact = False       # stiff variant: score every step that left the fast ranges
        self.block_threads = 0         # threads per block of the one-thread-per-chain kernels (0 = auto)
        self.chain_groups = 0          # pooled adaptation: launches per interval on the sampler's own streams (0 auto, 1 off)
        self.round_packing = 0         # d = 3: in-bounds proposals of a block packed at every solve round (0 auto, 1 off)
        self.rtol = 1e-6            # RateStateModel.py:374
        self.atol = 1e-10
        self.nmax = 500             # scipy dop853 nsteps default
        self.device = None          # torch device; None = current CUDA device
        self.last_status = None

    # -- helpers ---------------------------------------------------------
    def num_outputs(self) -> int:
        """``int(np.floor((t_final - t_start) / delta_t))`` (RateStateModel.py:358, quirk q8)."""
        return int(np.floor((self.t_final - self.t_start) / self.delta_t))

    def to_cfg(self) -> "_lib.RsfmCfg":
        cfg = _lib.default_cfg()
        cfg.a, cfg.b, cfg.mu_ref, cfg.V_ref, cfg.k1 = (float(self.a), float(self.b), float(self.mu_ref),
                                                      float(self.V_ref), float(self.k1))
        cfg.t_start, cfg.t_final, cfg.delta_t = float(self.t_start), float(self.t_final), float(self.delta_t)
        cfg.mu_t_zero = float(self.mu_t_zero)
        cfg.vstep_period, cfg.vstep_factor = float(self.vstep_period), float(self.vstep_factor)
        cfg.rtol, cfg.atol, cfg.nmax = float(self.rtol), float(self.atol), int(self.nmax)
        cfg.n_out = self.num_outputs()
        cfg.radiation_damping = 1 if self.RadiationDamping else 0
        cfg.loading = _LOADING[self.loading]
        cfg.integ_mode = _INTEG[self.integ_mode]
        cfg.observable = _OBSERVABLE[self.observable]
        cfg.solver_variant = _VARIANT[self.solver_variant]
        cfg.stiff_exact = 1 if self.stiff_exact else 0
        cfg.block_threads = int(self.block_threads)
        cfg.chain_groups = int(getattr(self, "chain_groups", 0))
        cfg.round_packing = int(getattr(self, "round_packing", 0))
        cfg.state_law = _LAW[self.state_law]
        if self.loading == "table":
            # the table lives on the device next to the model (the library reads it through the pointer)
            torch = _lib.require_cuda()
            tab = np.ascontiguousarray(self.load_table, dtype=np.float64).reshape(-1)
            if tab.size < 2:
                raise ValueError("load_table needs at least two entries")
            dev = self._device(torch)
            cur = self._load_table_dev
            if cur is None or cur[0].device != dev or cur[1].shape != tab.shape or not np.array_equal(cur[1], tab):
                self._load_table_dev = (torch.from_numpy(tab.copy()).to(dev), tab.copy())
            cfg.load_table_dev = self._load_table_dev[0].data_ptr()
            cfg.n_load_table = int(tab.size)
            cfg.load_dt = float(self.delta_t if self.load_dt is None else self.load_dt)
        return cfg

    def _device(self, torch):
        return torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())

    # -- batched forward solve (extension) -------------------------------
    def evaluate_batch(self, dc, a=None, b=None, data=None, want_acc=True, want_t=False, k1=None):
        """Solve for many parameter sets in one launch.

        dc, a, b: 1-D array-likes / CUDA tensors of equal length C (a, b optional).
        k1 (extension): a 1-D array of radiation-damping coefficients (RateStateModel.py:171, 351); the
        batch axis is then k1, ``dc`` must be ONE value (the common Dc) and a, b stay the model's scalars.
        Returns a dict of CUDA tensors: ``acc`` [n_out, C] (time-major; the observable: acceleration, or
        mu when ``self.observable == "mu"``), ``t``,
        ``sse`` [C] (when ``data`` is given), ``status``, ``filled``, ``nrhs``, ``nstep``.
        """
        torch = _lib.require_cuda()
        dev = self._device(torch)
        lib = _lib.load()

        def as_dev(x):
            if x is None:
                return None
            return torch.as_tensor(x, dtype=torch.float64).to(dev).contiguous().reshape(-1)

        dc_t, a_t, b_t, data_t = as_dev(dc), as_dev(a), as_dev(b), as_dev(data)
        cfg = self.to_cfg()
        if k1 is not None:
            if dc_t.numel() != 1 or a_t is not None or b_t is not None:
                raise ValueError("with k1 given, dc must be one value and a, b the model's scalars")
            cfg.sampled_param = _lib.PARAM_K1
            cfg.dc_fixed = float(dc_t.item())
            dc_t = as_dev(k1)                       # the library's per-chain scalar is k1 now (rsfm.h: sampled_param)
        cn = dc_t.numel()
        n_out = self.num_outputs()
        if data_t is not None and data_t.numel() != n_out:
            raise ValueError(f"data has {data_t.numel()} points, the model produces {n_out}")
        for name, x in (("a", a_t), ("b", b_t)):
            if x is not None and x.numel() != cn:
                raise ValueError(f"{name} must have the same length as dc")
        out = {
            "acc": torch.empty((n_out, cn), dtype=torch.float64, device=dev) if want_acc else None,
            "t": torch.empty((n_out, cn), dtype=torch.float64, device=dev) if want_t else None,
            "sse": torch.empty(cn, dtype=torch.float64, device=dev) if data_t is not None else None,
            "status": torch.empty(cn, dtype=torch.int32, device=dev),
            "filled": torch.empty(cn, dtype=torch.int32, device=dev),
            "nrhs": torch.empty(cn, dtype=torch.int64, device=dev),
            "nstep": torch.empty(cn, dtype=torch.int64, device=dev),
        }
        with torch.cuda.device(dev):
            rc = lib.rsfm_forward_batch(C.byref(cfg), cn, _lib.ptr(dc_t), _lib.ptr(a_t), _lib.ptr(b_t),
                                        _lib.ptr(data_t), _lib.ptr(out["acc"]), _lib.ptr(out["t"]),
                                        _lib.ptr(out["sse"]), _lib.ptr(out["status"]), _lib.ptr(out["filled"]),
                                        _lib.ptr(out["nrhs"]), _lib.ptr(out["nstep"]),
                                        _lib.current_stream(torch, dev))
        _lib.check(rc, "rsfm_forward_batch")
        return out

    # -- reference API -----------------------------------------------------
    def evaluate(self):
        """Reference ``evaluate()`` (RateStateModel.py:188-395): returns ``(t, acc, acc_noise)``.

        ``acc_noise = acc + |acc| * randn(N)`` draws from the global NumPy generator
        exactly like the reference (:392), so seeding with ``np.random.seed`` behaves
        the same.  An integrator failure leaves a zero tail (quirk q9) and, like
        SciPy, emits a ``UserWarning``.
        """
        if self.Dc is None:
            raise TypeError("RateStateModel.Dc must be set before evaluate()")   # reference: None / V_ref raises
        dc = float(np.ravel(np.asarray(self.Dc, dtype=np.float64))[0])
        out = self.evaluate_batch([dc], want_acc=True, want_t=True)
        acc = out["acc"][:, 0].cpu().numpy()
        t = out["t"][:, 0].cpu().numpy()
        status = int(out["status"][0].item())
        self.last_status = status
        if status != _lib.CHAIN_OK:
            import warnings
            msg = {_lib.CHAIN_NMAX: "larger nsteps is needed", _lib.CHAIN_HSMALL: "step size becomes too small"}
            warnings.warn(f"dop853: {msg.get(status, 'integration failed')}", UserWarning, stacklevel=2)
        # :392 -- for the reference observable acc[0] = 0, so |acc - acc[0]| is |acc| bit for bit; for
        # observable = "mu" the noise scales with the excursion from mu[0] = mu_ref in the same way
        acc_noise = acc + 1.0 * np.abs(acc - acc[0]) * np.random.randn(acc.shape[0])
        return t, acc, acc_noise
