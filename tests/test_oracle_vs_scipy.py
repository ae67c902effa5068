"""The oracle's DOP853 restatement against the third-party solver the reference
actually calls: scipy.integrate.ode('dop853') (RateStateModel.py:374).  Both are
driven with the SAME right-hand side (the oracle's C RHS through a callback), so
any difference is integrator logic.  CPU only; SciPy is in the image."""
import ctypes as C
import warnings

import numpy as np
import pytest
from scipy import integrate


def _call(orc, model, t0, y0, t1):
    lib = orc.lib()
    lib.orc_dop853_call.argtypes = [C.POINTER(orc.OrcModel), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.c_double, C.POINTER(orc.OrcStats)]
    lib.orc_dop853_call.restype = C.c_int
    st = orc.OrcStats()
    t = C.c_double(t0)
    y = (C.c_double * 3)(*y0)
    idid = lib.orc_dop853_call(C.byref(model), C.byref(t), y, t1, C.byref(st))
    return idid, t.value, np.array(list(y)), st


@pytest.mark.parametrize("dc,nint", [(1000.0, 60), (10.0, 60), (1.0, 40), (0.05, 12), (0.01, 6)])
def test_interval_by_interval_bit_exact(orc, dc, nint):
    model = orc.make_model(Dc=dc)
    calls = []

    def f(t, y):
        calls.append(t)
        return orc.rhs(model, t, y)

    r = integrate.ode(f).set_integrator("dop853", rtol=1e-6, atol=1e-10)
    r.set_initial_value([0.6, dc, 1.0], 0.0)
    for _ in range(nint):
        t0, y0 = r.t, np.array(r.y, dtype=float)
        calls.clear()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r.integrate(r.t + 0.1)
        idid, t1, y1, st = _call(orc, model, t0, y0, t0 + 0.1)
        assert idid == 1 and r.successful()
        assert t1 == r.t
        assert np.array_equal(y1, np.asarray(r.y)), (dc, t0)
        # SciPy may add benign repeat calls at step starts (stale irtrn); never fewer
        assert len(calls) >= st.nrhs and len(calls) - st.nrhs <= st.nstep + 1


def test_failure_semantics_nmax(orc):
    """Dc = 1e-4: more than nmax = 500 steps in the first interval -> idid -2 at the same t, y."""
    dc = 1e-4
    model = orc.make_model(Dc=dc)
    r = integrate.ode(lambda t, y: orc.rhs(model, t, y)).set_integrator("dop853", rtol=1e-6, atol=1e-10)
    r.set_initial_value([0.6, dc, 1.0], 0.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r.integrate(0.1)
    idid, t1, y1, st = _call(orc, model, 0.0, [0.6, dc, 1.0], 0.1)
    assert not r.successful() and idid == -2
    assert t1 == r.t and np.array_equal(y1, np.asarray(r.y))
    assert st.nstep == 501


def test_scipy_reject_rule_is_constant_shrink():
    """SciPy 1.18's C dop853 shrinks a rejected step by exactly 0.3 whatever err is (dop853.f
    would use h/min(1/0.3, err^(1/8)/0.9)).  The oracle and the CUDA kernel follow SciPy; this
    test documents the observation they rely on."""
    c2 = 0.05260015195876773
    lam = 100.0
    calls = []

    def f(t, y):
        calls.append(t)
        return -lam * np.asarray(y) + np.sin(t)

    r = integrate.ode(f).set_integrator("dop853", rtol=1e-6, atol=1e-10, first_step=0.011)
    r.set_initial_value([1.0], 0.0)
    r.integrate(0.5)
    c = np.array(calls)
    # attempt 1: the first non-zero call time is stage 2 (c2*h1) and stage 12 is called at h1; the
    # first non-zero call after that is stage 2 of the retry (c2*h2).  (SciPy may insert repeat
    # calls at t = 0.)
    nz = np.nonzero(c)[0]
    h1 = c[nz[0]] / c2
    i12 = int(np.nonzero(np.isclose(c, h1, rtol=1e-12, atol=0))[0][0])
    after = c[i12 + 1:]
    h2 = after[np.nonzero(after)[0][0]] / c2
    assert h1 == pytest.approx(0.011, rel=1e-12)
    assert h2 / h1 == pytest.approx(0.3, rel=1e-12)


@pytest.mark.parametrize("dc,n,t_end,period,factor", [(200.0, 300, 30.0, 10.0, 3.0), (2.0, 300, 30.0, 10.0, 3.0),
                                                       (0.05, 120, 12.0, 5.0, 10.0)])
def test_vstep_loading_pinned_on_scipy(orc, dc, n, t_end, period, factor):
    """The VSTEP extension has no reference counterpart: its CPU oracle is 'the same SciPy driver
    with only the loading line swapped' (RateStateModel.py:327-329).  oracle/scipy_port.py is that driver --
    scipy.integrate.ode('dop853') + the reference's Python RHS with the one line replaced -- and the C oracle's
    velocity-step trajectories (RHS and DOP853 restated) must reproduce it: same accepted steps, so agreement
    is at the level of libm-vs-NumPy rounding, far inside the 1e-9 trajectory gate."""
    from oracle import scipy_port
    pm = scipy_port.PortModel(number_time_steps=n, end_time=t_end)
    pm.loading, pm.vstep_period, pm.vstep_factor = "vstep", period, factor
    pm.Dc = dc
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t_p, acc_p, _ = pm.evaluate()
    om = orc.make_model(Dc=dc, number_time_steps=n, end_time=t_end, loading=orc.LOAD_VSTEP, vstep_period=period,
                        vstep_factor=factor)
    t_o, acc_o, st = orc.forward(om)
    assert st.filled == n and np.array_equal(t_o, t_p)
    scale = np.max(np.abs(acc_p))
    assert scale > 0
    # non-stiff: rounding level.  Stiff (Dc = 0.05): the step sequences of two correct solvers part at the first
    # differently rounded exp / log, and the trajectories then agree to the solver's tolerance only -- the gate is
    # a multiple of the oracle's own one-ulp reproducibility floor (conftest.oracle_ulp_floor)
    from conftest import oracle_ulp_floor
    _, _, floor = oracle_ulp_floor(orc, dc, number_time_steps=n, end_time=t_end, loading=orc.LOAD_VSTEP,
                                   vstep_period=period, vstep_factor=factor)
    assert (floor < 1e-11) == (dc >= 2.0)
    assert np.max(np.abs(acc_o - acc_p)) <= max(1e-11, 4.0 * floor) * scale
    # the loading is felt: nothing moves before the first velocity step, a transient right after it
    k_step = int(round(period / (t_end / n)))
    assert np.max(np.abs(acc_p[k_step:k_step + 20])) > 0.5 * scale > 1e3 * np.max(np.abs(acc_p[:k_step - 1]))
    # same number of RHS evaluations as SciPy made through the Python callback (benign repeats aside)
    if dc >= 2.0:
        assert pm.n_rhs >= st.nrhs and pm.n_rhs - st.nrhs <= st.nstep + n
    else:
        assert abs(pm.n_rhs - st.nrhs) <= 0.02 * st.nrhs


def test_mu_observable_pinned_on_scipy(orc):
    """The friction series mu[k] (RateStateModel.py:367, 385) of the C oracle against the SciPy driver."""
    from oracle import scipy_port
    pm = scipy_port.PortModel()
    pm.observable = "mu"
    pm.Dc = 1350.0
    _, mu_p, _ = pm.evaluate()
    _, mu_o, _, _, _, _ = orc.forward(orc.make_model(Dc=1350.0), full=True)
    assert mu_p[0] == 0.6 and np.max(np.abs(mu_o - mu_p)) <= 1e-15
    # forward_batch / chain replay score the same series when the model says so
    _, obs, _ = orc.forward_batch(orc.make_model(observable=orc.OBS_MU), [1350.0], want_acc=True)
    assert np.array_equal(obs[0], mu_o)


def _table(n, dt):
    tt = np.arange(n + 1) * dt
    return 0.5 * np.sin(0.7 * tt) * np.exp(-tt / 30.0) + 0.3 * (tt > 20.0)


@pytest.mark.parametrize("dc", [300.0, 1350.0])
def test_slip_law_and_tabulated_loading_pinned_on_scipy(orc, dc):
    """Extensions of SURVEY 8f.4, each ONE line of the reference's RHS swapped in the SciPy driver
    (RateStateModel.py:340 for Ruina's slip law, :327-329 for a piecewise-linear tabulated load): the C oracle
    reproduces SciPy's trajectories to the last bit, as it does for the reference's own RHS."""
    from oracle import scipy_port
    tab = _table(500, 0.1)
    for law, loading in (("slip", "sine_decay"), ("aging", "table"), ("slip", "table")):
        pm = scipy_port.PortModel()
        pm.Dc, pm.state_law, pm.loading, pm.load_table, pm.load_dt = dc, law, loading, tab, 0.1
        _, acc_p, _ = pm.evaluate()
        kw = dict(state_law=orc.LAW_SLIP if law == "slip" else orc.LAW_AGING)
        if loading == "table":
            kw.update(loading=orc.LOAD_TABLE, load_table=tab, load_dt=0.1)
        t_o, acc_o, st = orc.forward(orc.make_model(Dc=dc, **kw))
        assert st.istate == 1 and st.filled == 500
        assert np.max(np.abs(acc_o - acc_p)) <= 1e-13 * np.max(np.abs(acc_p)), (law, loading)
    # the swapped lines are felt: the slip law moves the trajectory at the 1e-7 .. 1e-5 level (both laws share
    # their linearisation about steady state), the tabulated load changes it altogether
    base = orc.forward(orc.make_model(Dc=dc))[1]
    slip = orc.forward(orc.make_model(Dc=dc, state_law=orc.LAW_SLIP))[1]
    tabd = orc.forward(orc.make_model(Dc=dc, loading=orc.LOAD_TABLE, load_table=tab, load_dt=0.1))[1]
    scale = np.max(np.abs(base))
    assert 1e-8 * scale < np.max(np.abs(slip - base)) < 1e-3 * scale
    assert np.max(np.abs(tabd - base)) > 0.1 * scale
    # beyond the ends the table is held constant
    short = orc.forward(orc.make_model(Dc=dc, loading=orc.LOAD_TABLE, load_table=np.append(tab[:201], tab[200]), load_dt=0.1))[1]
    const = orc.forward(orc.make_model(Dc=dc, loading=orc.LOAD_TABLE, load_table=np.concatenate([tab[:201], np.full(300, tab[200])]),
                                       load_dt=0.1))[1]
    assert np.array_equal(short, const)
