"""Premises of the stiff kernel variant, checked on the CPU oracle step by step.

The variant takes a DOP853 trial step as rejected, without scoring it, when the step starts inside half of the fast
step's series ranges and one of its stages leaves them (f' or A', in the frame re-based on the current load level).
SciPy's own rule is err <= 1; the two agree if the exact arithmetic never accepts such a step.  The oracle's step hook
reports, for every attempted step, err and the excess of the series arguments over the limits at the start and at
the worst stage."""
import ctypes as C

import numpy as np
import pytest


def _steps(orc, **kw):
    rec = []
    hook_t = C.CFUNCTYPE(None, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double)
    cb = hook_t(lambda x, h, err, start, worst: rec.append((x, h, err, start, worst)))
    lib = orc.lib()
    lib.orc_set_step_hook.argtypes = [hook_t]
    lib.orc_set_step_hook(cb)
    try:
        _, _, st = orc.forward(orc.make_model(**kw))
    finally:
        lib.orc_set_step_hook(hook_t())
    assert st.istate == 1 and len(rec) == st.nstep
    return np.array(rec)


@pytest.mark.parametrize("dc, n, period, factor", [(0.05, 600, 20.0, 10.0), (0.03, 300, 10.0, 10.0),
                                                   (0.15, 600, 20.0, 10.0), (0.05, 600, 15.0, 3.0)])
def test_steps_leaving_the_fast_ranges_from_inside_are_never_accepted(orc, dc, n, period, factor):
    r = _steps(orc, Dc=dc, number_time_steps=n, end_time=0.1 * n, loading=orc.LOAD_VSTEP, vstep_period=period,
               vstep_factor=factor)
    err, start, worst = r[:, 2], r[:, 3], r[:, 4]
    wild = (start < 0.5) & ~(worst < 1.0)                  # starts inside half the ranges, a stage leaves them
    accepted = err <= 1.0                                  # NaN compares false: rejected, as in dop853
    assert wild.sum() > 100                                # the regime is the stiff one: such steps are common ...
    assert 0.05 < wild.mean() < 0.35
    assert not np.any(wild & accepted)                     # ... and the exact arithmetic rejects every one of them
    # the steps the variant scores with the fast step are the overwhelming majority
    assert np.mean((start < 0.5) & (worst < 1.0)) > 0.6
    # steps that START outside the ranges (general-range step in the variant) are the transients after a jump
    assert np.mean(start >= 0.5) < 0.05
