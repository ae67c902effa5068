"""Device RHS alone (unit level): rsf_rhs in its two forms against values of the reference's
own nested friction(t, y) (tests/golden/rhs_values.json, captured unmodified by oracle/make_golden.py)."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, rhs_term_scales

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("general", [0, 1])
def test_device_rhs_matches_reference_friction(cuda, pkg, general):
    """general = 0: a single RHS evaluation exactly as the solver makes it (cancellation-free short series near
    sliding steady state, falling back to the reference's formulas outside their range); general = 1: the
    reference's formulas at every state.  Gate: 1e-15 of the magnitude of the terms each component is a
    difference of (the reference's own rounding noise in those differences), both damping modes."""
    torch = cuda
    lib = pkg._lib.load()
    g = load_golden("rhs_values.json")
    rows = g["rows"]
    worst = 0.0
    for damping in (1.0, 0.0):
        sel = rows[rows[:, 0] == damping]
        m = pkg.RateStateModel()
        m.RadiationDamping = bool(damping)
        cfg = m.to_cfg()
        n = sel.shape[0]
        dev = lambda col: torch.from_numpy(np.ascontiguousarray(sel[:, col])).cuda()
        t, mu, th, dc = dev(2), dev(3), dev(4), dev(1)
        out = torch.empty((3, n), dtype=torch.float64, device="cuda")
        pkg._lib.check(lib.rsfm_rhs_eval(C.byref(cfg), n, t.data_ptr(), mu.data_ptr(), th.data_ptr(), dc.data_ptr(),
                                         None, None, general, out.data_ptr(), None), "rsfm_rhs_eval")
        got = out.cpu().numpy().T                                    # [n, 3] = (mu', theta', V')
        for i in range(n):
            worst = max(worst, float(np.max(np.abs(got[i] - sel[i, 6:9]) / rhs_term_scales(sel[i]))))
    assert worst <= 1e-15, worst


def test_device_rhs_fast_and_general_forms_agree(cuda, pkg):
    """Inside the fast ranges the short-series form and the reference's formulas are the same function; the
    cancellation-free form is the more accurate of the two (checked against 80-bit long double)."""
    torch = cuda
    lib = pkg._lib.load()
    rng = np.random.default_rng(1)
    n = 4096
    dc = rng.uniform(130.0, 9000.0, n)
    th = dc * (1 + rng.uniform(-1.5e-3, 1.5e-3, n) / (1 + 0.014 / 0.011))
    mu = 0.6 + rng.uniform(-1e-4, 1e-4, n)
    t = rng.uniform(0, 50, n)
    cfg = pkg.RateStateModel().to_cfg()
    args = [torch.from_numpy(x).cuda() for x in (t, mu, th, dc)]
    outs = []
    for general in (0, 1):
        out = torch.empty((3, n), dtype=torch.float64, device="cuda")
        pkg._lib.check(lib.rsfm_rhs_eval(C.byref(cfg), n, *[a.data_ptr() for a in args], None, None, general,
                                         out.data_ptr(), None))
        outs.append(out.cpu().numpy())
    L = np.longdouble
    a, b, kp = L(0.011), L(0.014), L(0.1) / dc.astype(L)
    mul, thl, dcl = mu.astype(L), th.astype(L), dc.astype(L)
    v = np.exp((mul - L(0.6) - b * np.log(thl / dcl)) / a)
    dth = 1 - v * thl / dcl
    # the ARGUMENTS of the load term are formed in double, as the reference forms them (RateStateModel.py:327-329:
    # np.exp(-t/a1) * np.sin(a2*t)); 10 t carries up to 5e-14 of rounding at t = 50, which is not the RHS's error
    dmu = kp * (1 + np.exp((-t / 20.0).astype(L)) * np.sin((10.0 * t).astype(L))) - kp * v
    dv = v / a * (dmu - b / thl * dth)
    dmu = dmu - L(1e-7) * dv
    dv = v / a * (dmu - b / thl * dth)
    truth = np.array([dmu, dth, dv], dtype=L)
    scale = np.array([kp * 2, np.full(n, 2.0), v / a * (kp * 2 + b / thl * 2)], dtype=L)
    err_fast = np.max(np.abs(outs[0] - truth) / scale)
    err_gen = np.max(np.abs(outs[1] - truth) / scale)
    assert err_gen <= 2e-15 and err_fast <= 1e-15, (err_fast, err_gen)
    assert err_fast <= max(err_gen, 3e-16), (err_fast, err_gen)
