"""``MCMC`` -- drop-in for the reference Metropolis sampler, fused on B200.

Mirrors the reference class (MCMC.py:4-544): same constructor arguments, the same
attributes after ``sample()`` (``std2``, ``Vstart``, ``nburn`` ...), the same return
shape ``(d, nsamples + 1 - nburn)`` and the same per-iteration semantics (SURVEY.md
Appendix A), including the quirks that depend on the prior container type:

  * ``["Uniform", lo, hi]`` (main.py:54): ``len(qpriors) = 3`` degrees-of-freedom
    correction (q5) and NO adaptation, because ``qpriors.keys()`` raises inside a
    bare ``except`` (q2);
  * ``{1: lo, 2: hi}`` (MCMC.py:38): ``len = 2`` and the per-chain, last-10-samples,
    Cholesky-factor-used-as-covariance update of MCMC.py:200-204 (q3).

The loop itself (propose, bounds test, forward solve, SSE, accept, sigma^2 draw,
adaptation) runs inside one CUDA kernel (``rsf_mcmc_kernel`` in librsfm) for all
chains and all iterations; nothing is sampled or integrated on the CPU.

Extensions (keyword-only; defaults reproduce the reference): ``n_chains``,
``seed``, ``device``, ``param_names`` (("Dc",), ("a", "b", "Dc") or ("k1",): the radiation-damping
coefficient with ``model.Dc`` held fixed), ``bounds``,
``deterministic_inputs``, ``compat_adapt``, ``adapt`` ("pooled" = Haario-style
covariance pooled over chains and ranks), ``shard`` (split chains over
torch.distributed ranks).
"""
import ctypes as C
import time

import numpy as np

from . import _lib
from .sharding import ChainShard, rank_and_world

# result arrays larger than this are copied to the host in chunks while the chains are still running
_OVERLAP_BYTES = 128 << 20
_CHUNK_BYTES = 96 << 20


class _ResultPipe:
    """Device -> host transfer of the result arrays, overlapped with the iterations.

    The device arrays are time-major (``[rows, ...]``); rows become final as iterations complete, so
    finished row ranges are copied on a side stream into page-locked host buffers of the same layout
    while the next launch runs (torch's caching host allocator recycles the buffers between calls;
    pageable destinations cost ~8 ms per call for the 18 MB of cfg 2 against < 1 ms pinned).  The
    reference-shaped ``[C, d, n]`` outputs are transposed VIEWS of those buffers: no second copy."""

    def __init__(self, torch, dev, arrays):
        self.torch, self.dev = torch, dev
        self.stream = torch.cuda.Stream(dev)
        self.items = []                               # (device tensor, first row wanted, host tensor)
        for t, first in arrays:
            shape = (t.shape[0] - first,) + tuple(t.shape[1:])
            try:
                h = torch.empty(shape, dtype=t.dtype, pin_memory=True)
            except RuntimeError:                      # locked-memory limit reached: plain pageable copy
                h = torch.empty(shape, dtype=t.dtype)
            self.items.append((t, first, h, [first]))

    def push(self, upto_rows):
        """Rows below ``upto_rows[i]`` of array i are final once the work queued so far on the current
        stream is done: copy what has not been copied yet."""
        torch = self.torch
        ev = torch.cuda.current_stream(self.dev).record_event()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            for (t, first, h, done), upto in zip(self.items, upto_rows):
                lo, hi = done[0], min(int(upto), t.shape[0])
                if hi > lo:
                    h[lo - first:hi - first].copy_(t[lo:hi], non_blocking=True)
                    done[0] = hi

    def finish(self):
        self.push([t.shape[0] for t, _, _, _ in self.items])
        self.stream.synchronize()
        return [h.numpy() for _, _, h, _ in self.items]


class MCMC:
    def __init__(self, model, data, dc_true, qpriors, qstart, nsamples=100, lstm_model={},
                 adapt_interval=10, verbose=True, *, n_chains=1, seed=None, device=None,
                 param_names=("Dc",), bounds=None, deterministic_inputs=None, compat_adapt=None,
                 adapt=None, adapt_start=100, shard=False, chain_id0=0, spec_depth=0, resume=None):
        # reference attributes, MCMC.py:88-99
        self.model = model
        self.qstart = qstart
        self.qpriors = qpriors
        self.nsamples = nsamples
        self.nburn = int(nsamples / 2)
        self.verbose = verbose
        self.adapt_interval = adapt_interval
        self.data = data
        self.lstm_model = lstm_model
        self.n0 = 0.01
        self.qstart_limits = np.array([[self.qpriors[1], self.qpriors[2]]])
        self.dc_true = dc_true
        # extensions
        self.n_chains = int(n_chains)
        self.seed = seed
        self.device = device
        self.param_names = tuple(param_names)
        self.bounds = bounds
        self.deterministic_inputs = deterministic_inputs
        self.adapt = adapt
        self.adapt_start = int(adapt_start)
        self.shard = bool(shard)
        self.chain_id0 = int(chain_id0)        # global id of chain 0 when not sharding (Philox counter)
        self.spec_depth = int(spec_depth)      # speculation tree depth: 0 auto, 1 off, 2..5 forced
        self.resume = resume                   # checkpoint dict / JSON file from MCMC.checkpoint()
        if self.param_names not in (("Dc",), ("a", "b", "Dc"), ("k1",)):
            raise ValueError("param_names must be ('Dc',), ('a', 'b', 'Dc') or ('k1',)")
        if compat_adapt is None:
            # reference behaviour follows the container type of qpriors (q2 / q3)
            compat_adapt = hasattr(qpriors, "keys") and adapt is None
        self.compat_adapt = bool(compat_adapt)
        if self.compat_adapt and len(self.param_names) != 1:
            raise ValueError("compat_adapt reproduces the reference's d = 1 update only")
        if adapt not in (None, "pooled"):
            raise ValueError("adapt must be None or 'pooled'")
        # results
        self.std2 = None
        self.Vstart = None
        self.acceptance_ratio = None
        self.samples_device = None
        self.stats = {}

    # ------------------------------------------------------------------
    def _bounds(self, d):
        if self.bounds is not None:
            b = np.asarray(self.bounds, dtype=np.float64).reshape(d, 2)
            return b[:, 0].copy(), b[:, 1].copy()
        # the reference has ONE shared bound pair (MCMC.py:98, q10)
        lo = np.full(d, float(self.qpriors[1]))
        hi = np.full(d, float(self.qpriors[2]))
        return lo, hi

    def _start_values(self, torch, dev, d, shard):
        """q0 as a [d, C_local] device tensor from a scalar, [d], [C] or [C, d] qstart."""
        q = np.asarray(self.qstart, dtype=np.float64)
        cg = self.n_chains
        if q.ndim == 0:
            q0 = np.full((d, cg), float(q))
        elif q.ndim == 1 and q.size == d and (d > 1 or cg == 1):
            q0 = np.repeat(q.reshape(d, 1), cg, axis=1)
        elif q.ndim == 1 and d == 1 and q.size == cg:
            q0 = q.reshape(1, cg)
        elif q.ndim == 2 and q.shape == (cg, d):
            q0 = np.ascontiguousarray(q.T)
        else:
            raise ValueError(f"qstart of shape {q.shape} does not match d = {d}, n_chains = {cg}")
        q0 = q0[:, shard.start:shard.stop]
        return torch.from_numpy(np.ascontiguousarray(q0)).to(dev)

    # ------------------------------------------------------------------
    def sample(self, MAKE_ANIMATIONS=False):
        """Run the chains.  Returns ``qparams[:, nburn:]`` like MCMC.py:544.

        Shape ``(d, nsamples + 1 - nburn)`` for one chain, ``(n_chains_local, d,
        nsamples + 1 - nburn)`` otherwise.  ``MAKE_ANIMATIONS`` (matplotlib/ffmpeg
        movie, MCMC.py:471-492) is outside the scope of this package and ignored.
        """
        if self.lstm_model:
            raise NotImplementedError("the reduced-order-model branch (MCMC.py:124-125) is dead code in the "
                                      "reference (no such attribute exists) and is not provided")
        if not hasattr(self.model, "to_cfg"):
            raise TypeError("model must be a bayesian-markov-chain-monte-carlo_b200 RateStateModel: the forward "
                            "solve runs inside the CUDA kernel, a Python evaluate() cannot be called from it")
        torch = _lib.require_cuda()
        lib = _lib.load()
        dev = torch.device(self.device) if self.device is not None else self.model._device(torch)
        d = len(self.param_names)
        shard = ChainShard.for_current_rank(self.n_chains) if self.shard else ChainShard(0, self.n_chains, self.n_chains)
        cl = shard.count
        id0 = shard.start if self.shard else self.chain_id0
        n_out = self.model.num_outputs()
        data = np.ascontiguousarray(np.asarray(self.data, dtype=np.float64).reshape(-1))
        if data.size != n_out:
            raise ValueError(f"data has {data.size} points, the model produces {n_out}")

        cfg = self.model.to_cfg()
        cfg.n_params = d
        sample_k1 = self.param_names == ("k1",)
        if sample_k1:
            # the chain's scalar is model.k1 (RateStateModel.py:171, 351) instead of model.Dc (MCMC.py:381);
            # Dc is what the caller left on the model
            if self.model.Dc is None:
                raise TypeError("param_names=('k1',) needs model.Dc set (the fixed critical slip distance)")
            cfg.sampled_param = _lib.PARAM_K1
            cfg.dc_fixed = float(np.ravel(np.asarray(self.model.Dc, dtype=np.float64))[0])
        lo, hi = self._bounds(d)
        for j in range(d):
            cfg.lo[j], cfg.hi[j] = lo[j], hi[j]
        cfg.n0 = float(self.n0)
        cfg.n_prior_len = len(self.qpriors)                       # MCMC.py:261 (q5)
        cfg.adapt_interval = int(self.adapt_interval)
        cfg.spec_depth = self.spec_depth
        cfg.adapt_mode = (_lib.ADAPT_COMPAT if self.compat_adapt else
                          _lib.ADAPT_POOLED if self.adapt == "pooled" else _lib.ADAPT_NONE)
        seed = self.seed
        if self.resume is not None:
            if isinstance(self.resume, str):
                from .ndarray_json import load_object
                self.resume = load_object(self.resume)
            if self.resume is None:
                raise ValueError("checkpoint could not be read")
            seed = int(self.resume["seed"])            # the Philox key and chain ids are part of the state
            id0 = int(self.resume["chain_id0"])
        if seed is None:
            # the reference draws from the global NumPy generator; derive the Philox key from it so
            # that np.random.seed(...) before the call makes runs reproducible here too
            s = np.random.randint(0, 2 ** 31 - 1, size=2)
            seed = (int(s[0]) << 31) | int(s[1])
        self.seed_used = int(seed)
        self._pooled_stats, self._pooled_state = {}, None

        t_begin = time.perf_counter()
        with torch.cuda.device(dev):
            stream = _lib.current_stream(torch, dev)
            data_t = torch.from_numpy(data).to(dev)
            q0 = self._start_values(torch, dev, d, shard)
            handle = lib.rsfm_create(C.byref(cfg), cl, C.c_uint64(self.seed_used), C.c_uint64(id0))
            if not handle:
                _lib.check(-1, "rsfm_create")
            try:
                _lib.check(lib.rsfm_init(handle, _lib.ptr(q0), _lib.ptr(data_t), stream), "rsfm_init")
                s2_0 = torch.empty(cl, dtype=torch.float64, device=dev)
                chol0 = torch.empty((d * (d + 1) // 2, cl), dtype=torch.float64, device=dev)
                _lib.check(lib.rsfm_get_state(handle, None, None, _lib.ptr(s2_0), _lib.ptr(chol0), None, None,
                                              None, None, stream), "rsfm_get_state")
                if self.resume is not None:
                    q0, s2_0 = self._apply_checkpoint(torch, lib, handle, dev, d, cl, stream)
                ns = int(self.nsamples)
                nb = self.nburn
                chain = torch.empty((ns + 1, d, cl), dtype=torch.float64, device=dev)
                std2 = torch.empty((ns + 1, cl), dtype=torch.float64, device=dev)
                accept = torch.empty((ns, cl), dtype=torch.uint8, device=dev)
                want_draws = cl == 1 and self.deterministic_inputs is None
                draws = torch.empty((ns, d + 2, cl), dtype=torch.float64, device=dev) if want_draws else None
                chain[0] = q0
                std2[0] = s2_0
                pipe = _ResultPipe(torch, dev, [(chain, nb), (std2, nb), (accept, 0)])
                out_bytes = (ns + 1 - nb) * (d + 1) * cl * 8 + ns * cl

                def run_iters(k, done):
                    _lib.check(lib.rsfm_run(handle, k, _lib.ptr(chain[1 + done:]), _lib.ptr(std2[1 + done:]),
                                            _lib.ptr(accept[done:]), _lib.ptr(draws[done:]) if draws is not None else None,
                                            stream), "rsfm_run")

                if self.deterministic_inputs is not None:
                    self._run_deterministic(torch, lib, handle, dev, d, cl, ns, chain, std2, accept, stream)
                elif cfg.adapt_mode == _lib.ADAPT_POOLED:
                    self._run_pooled(torch, lib, handle, dev, d, cl, ns, run_iters, pipe, stream)
                elif out_bytes <= _OVERLAP_BYTES or ns < 2:
                    run_iters(ns, 0)
                else:
                    # large outputs (cfg 5: 0.5 GB per GPU): burn-in in one launch, then launches of a few dozen
                    # iterations whose results go to the host while the next launch runs.  The chains do not
                    # depend on the launch partition (draws are keyed by chain and iteration).
                    per_iter = (d + 1) * cl * 8 + cl
                    step = max(20, _CHUNK_BYTES // per_iter)
                    done = 0
                    if nb > step:
                        run_iters(nb - 1, 0)
                        done = nb - 1
                        pipe.push([1 + done, 1 + done, done])
                    while done < ns:
                        k = min(step, ns - done)
                        run_iters(k, done)
                        done += k
                        pipe.push([1 + done, 1 + done, done])
                acc_cnt = torch.empty(cl, dtype=torch.int32, device=dev)
                status = torch.empty(cl, dtype=torch.int32, device=dev)
                nrhs = torch.empty(cl, dtype=torch.int64, device=dev)
                nstep = torch.empty(cl, dtype=torch.int64, device=dev)
                _lib.check(lib.rsfm_get_state(handle, None, None, None, None, _lib.ptr(acc_cnt), _lib.ptr(status),
                                              _lib.ptr(nrhs), _lib.ptr(nstep), stream), "rsfm_get_state")
                st_q = torch.empty((d, cl), dtype=torch.float64, device=dev)
                st_sse = torch.empty(cl, dtype=torch.float64, device=dev)
                st_s2 = torch.empty(cl, dtype=torch.float64, device=dev)
                st_chol = torch.empty((d * (d + 1) // 2, cl), dtype=torch.float64, device=dev)
                _lib.check(lib.rsfm_get_state(handle, _lib.ptr(st_q), _lib.ptr(st_sse), _lib.ptr(st_s2),
                                              _lib.ptr(st_chol), None, None, None, None, stream), "rsfm_get_state")
                st_ring = None
                if cfg.adapt_mode == _lib.ADAPT_COMPAT:
                    st_ring = torch.empty((int(self.adapt_interval), cl), dtype=torch.float64, device=dev)
                    _lib.check(lib.rsfm_get_ring(handle, _lib.ptr(st_ring), stream), "rsfm_get_ring")
                self._final_state = (st_q, st_sse, st_s2, st_chol, int(lib.rsfm_iteration(handle)), st_ring)
                # ---- reference-shaped host outputs: the last rows go to the host, everything is awaited ----
                chain_t, std2_t, accept_t = pipe.finish()                  # [n, d, C], [n, C], [ns, C]
                tot = (C.c_uint64 * 9)()
                _lib.check(lib.rsfm_get_totals(handle, tot, stream), "rsfm_get_totals")
                torch.cuda.synchronize(dev)
            finally:
                lib.rsfm_destroy(handle)
        elapsed = time.perf_counter() - t_begin

        self.samples_device = chain                       # [nsamples+1, d, C_local], start value included
        self.std2_device = std2
        self.accept_device = accept
        in_bounds = None
        chain_h, std2_h, accept_h = chain_t.transpose(2, 1, 0), std2_t.T, accept_t.T   # [C, d, n], [C, n], [C, ns]
        self.Vstart = self._vstart_host(chol0, d)
        self.status = status.cpu().numpy()
        n_acc = acc_cnt.cpu().numpy().astype(np.int64)
        self.acceptance_ratio = n_acc / float(self.nsamples)
        self.stats = {
            "elapsed_s": elapsed, "n_chains_local": cl, "chain_id0": id0,
            "nsolves": int(tot[0]), "nrhs": int(tot[1]), "nstep": int(tot[2]),
            "nsolves_stopped_early": int(tot[5]), "nsolves_executed": int(tot[6]),
            "nrhs_deciding": int(tot[7]), "nstep_deciding": int(tot[8]),
            "failed_chains": int((status != 0).sum().item()),
        }
        self.stats.update(getattr(self, "_pooled_stats", {}))
        if cl == 1:
            last = None
            if draws is not None:
                dr = draws[:, :, 0].cpu().numpy()
                in_bounds = ~np.isnan(dr[:, d])
                evaluated = dr[in_bounds, :d]
                if evaluated.shape[0]:
                    last = evaluated[-1, d - 1]
                if self.verbose:
                    for i in range(self.nsamples):               # MCMC.py:503-504
                        print(i, bool(accept_h[0, i]))
                        print("Generated Sample ---- ", dr[i, d - 1] if d == 1 else dr[i, :d])
                # the reference leaves model.Dc at the last evaluated proposal, a 1-element array (q6)
                left = np.array([last if last is not None else float(np.ravel(self.qstart)[-1])])
                if sample_k1:
                    self.model.k1 = float(left[0])
                else:
                    self.model.Dc = left
            if self.verbose:
                print("acceptance ratio:", self.acceptance_ratio[0])      # MCMC.py:530
            self.std2 = np.ascontiguousarray(std2_h[0])                    # MCMC.py:533
            self.accepts = np.ascontiguousarray(accept_h[0])
            return np.ascontiguousarray(chain_h[0])                        # (d, nsamples+1-nburn), MCMC.py:544
        if self.verbose:
            print("acceptance ratio:", float(self.acceptance_ratio.mean()))
        self.std2 = std2_h
        self.accepts = accept_h
        return chain_h

    # ------------------------------------------------------------------
    @staticmethod
    def _vstart_host(chol0, d):
        c = chol0[:, 0].cpu().numpy()
        if d == 1:
            return np.array([[c[0]]])                            # (1, 1) like MCMC.py:266
        low = np.zeros((d, d))
        low[np.tril_indices(d)] = c
        return low @ low.T

    def _run_deterministic(self, torch, lib, handle, dev, d, cl, ns, chain, std2, accept, stream):
        """Host-supplied randomness: step-for-step replay."""
        di = self.deterministic_inputs

        def dev3(x, shape):
            t = torch.as_tensor(np.asarray(x, dtype=np.float64)).reshape(shape).to(dev).contiguous()
            return t

        key = "z" if "z" in di else "proposals"
        prop = dev3(di[key], (ns, d, cl))
        uni = dev3(np.nan_to_num(np.asarray(di["uniforms"], dtype=np.float64), nan=0.5), (ns, cl))
        gam = dev3(di["gammas"], (ns, cl))
        _lib.check(lib.rsfm_run_deterministic(handle, ns, _lib.ptr(prop), 1 if key == "z" else 0, _lib.ptr(uni),
                                              _lib.ptr(gam), _lib.ptr(chain[1:]), _lib.ptr(std2[1:]),
                                              _lib.ptr(accept), stream), "rsfm_run_deterministic")
        torch.cuda.synchronize(dev)

    def _run_pooled(self, torch, lib, handle, dev, d, cl, ns, run_iters, pipe, stream):
        """Haario-style adaptive Metropolis with the covariance pooled over all chains of all ranks
        (generalises MCMC.py:162-204, 523-527), every step of it on the device.

        Per interval j of ``adapt_interval`` iterations, without any host synchronisation:
          main stream   [update(j-2)]  run(j)  partials(j)                    (rsfm_pooled_update / rsfm_run /
          side stream                          all-gather(j) ............      rsfm_pooled_partials)
        ``partials`` are sums over fixed groups of 1,024 chains aligned on the global chain id; the
        all-gather (NCCL; a device copy for one rank) puts every rank's rows in global chain order, and
        ``update`` adds them to the running moments in that order, forms (2.38^2/d) cov and its Cholesky
        factor in closed form and installs it for every chain -- the same bits on every rank and for any
        number of ranks.  The all-gather of interval j overlaps run(j+1); its factor is used from
        interval j+2 on (adaptation lags one interval, as SURVEY 8e allows)."""
        from .adaptation import PooledAdaptation
        w = max(1, int(self.adapt_interval))
        rank, world = rank_and_world() if self.shard else (0, 1)
        pool = PooledAdaptation(torch, lib, handle, dev, d, self.n_chains, world, self.adapt_start, stream)
        done = int(lib.rsfm_iteration(handle))            # > 0 after a resume
        if self.resume is not None and "pooled_moments" in self.resume:
            if done % w:
                raise ValueError("a pooled-adaptation checkpoint continues exactly only from an adaptation "
                                 f"boundary (iteration {done} is not a multiple of adapt_interval = {w})")
            pool.preload(self.resume["pooled_moments"], self.resume["pooled_pending"], done)
        pos = 0
        while pos < ns:
            k = min(w - (done % w), ns - pos)             # intervals end on absolute multiples of adapt_interval
            pool.before_interval()
            run_iters(k, pos)
            pos += k
            done += k
            pool.after_interval(done)
            pipe.push([1 + pos, 1 + pos, pos])
        self._pooled_state = pool.finish()
        self.adapt_history = pool.history
        self._pooled_stats = dict(pool.stats, chain_groups=int(lib.rsfm_chain_groups(handle)))

    # ------------------------------------------------------------------
    def checkpoint(self, filename=None):
        """State needed to continue the chains exactly where they stopped (the reference has no
        checkpoint): current q, SSE, sigma^2, proposal factor, iteration count, seed,
        and the adaptation state -- the per-chain sample ring of the reference's windowed update
        (compat mode), or the pooled moments plus the gathered rows of the last interval (pooled mode;
        exact from an adaptation boundary).  With counter-based Philox draws,
        ``MCMC(..., resume=ckpt).sample()`` continues the very same chains.  Returned as a dict of NumPy
        arrays and, if ``filename`` is given, written in the reference's JSON ndarray format."""
        if getattr(self, "_final_state", None) is None:
            raise RuntimeError("call sample() first")
        q, sse, s2, chol, iteration, ring = self._final_state
        ck = {"q": q.cpu().numpy(), "sse": sse.cpu().numpy(), "sigma2": s2.cpu().numpy(), "chol": chol.cpu().numpy(),
              "iteration": int(iteration), "seed": int(self.seed_used), "chain_id0": int(self.stats["chain_id0"]),
              "param_names": list(self.param_names), "n_chains_local": int(q.shape[1])}
        if ring is not None:
            ck["ring"] = ring.cpu().numpy()
        if self.adapt == "pooled" and getattr(self, "_pooled_state", None) is not None:
            ck["pooled_moments"] = self._pooled_state[0].cpu().numpy()
            ck["pooled_pending"] = self._pooled_state[1].cpu().numpy()
        if filename is not None:
            from .ndarray_json import save_object
            save_object(ck, filename)
        return ck

    def _apply_checkpoint(self, torch, lib, handle, dev, d, cl, stream):
        ck = self.resume
        if int(ck["n_chains_local"]) != cl or list(ck["param_names"]) != list(self.param_names):
            raise ValueError("checkpoint does not match this sampler (chains / parameters)")
        def dev_t(x, shape):
            return torch.as_tensor(np.asarray(x, dtype=np.float64)).reshape(shape).to(dev).contiguous()
        q = dev_t(ck["q"], (d, cl)); sse = dev_t(ck["sse"], (cl,)); s2 = dev_t(ck["sigma2"], (cl,))
        chol = dev_t(ck["chol"], (d * (d + 1) // 2, cl))
        _lib.check(lib.rsfm_set_state(handle, _lib.ptr(q), _lib.ptr(sse), _lib.ptr(s2), _lib.ptr(chol),
                                      int(ck["iteration"]), stream), "rsfm_set_state")
        if self.compat_adapt:
            if "ring" not in ck:
                raise ValueError("checkpoint holds no sample ring: it was not written by a compat-adaptation run")
            ring = dev_t(ck["ring"], (int(self.adapt_interval), cl))
            _lib.check(lib.rsfm_set_ring(handle, _lib.ptr(ring), stream), "rsfm_set_ring")
        return q, s2

    def diagnostics(self, max_lag=None):
        """Split-R-hat and bulk ESS of the post-burn-in draws, pooled over ranks."""
        from .diagnostics import chain_diagnostics
        if self.samples_device is None:
            raise RuntimeError("call sample() first")
        return chain_diagnostics(self.samples_device[self.nburn:], max_lag=max_lag)

    def save_samples(self, filename):
        """Write the post-burn-in samples in the reference's JSON ndarray format (json_save_load.py:37-38)."""
        from .ndarray_json import save_object
        if self.samples_device is None:
            raise RuntimeError("call sample() first")
        nb = self.nburn
        obj = {
            "param_names": list(self.param_names),
            "samples": self.samples_device[nb:].permute(2, 1, 0).contiguous().cpu().numpy(),
            "std2": self.std2_device[nb:].t().contiguous().cpu().numpy(),
            "acceptance_ratio": np.asarray(self.acceptance_ratio, dtype=np.float64),
            "nsamples": int(self.nsamples), "nburn": int(nb), "seed": int(self.seed_used),
        }
        save_object(obj, filename)
