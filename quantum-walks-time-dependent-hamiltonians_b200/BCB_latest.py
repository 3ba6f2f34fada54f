"""Drop-in for the reference's ``BCB_latest.py``: BCB with the thesis' production sweep
(75 beta values x 4 times {pi/4, 1, pi/2, 2} * sqrt(N); BCB_latest.py:242-264) and the
file names ``Robustness.py`` reads (BCB_latest.py:287-289)."""

import numpy as np

from . import _dropin

dimension = None
step_function = "lin"
topology = "circular"
type_of_qw = "dependent"
rtolerance = 1e-6
atolerance = 1e-6
n_steps = None
step_size = 1e-3
gamma_on = "oracle"
integrator = "rk4"           # 'rk45': SciPy's adaptive RK45 as the reference ships it
ti_method = "exact"         # time-independent comparator: 'exact' (Chebyshev) or 'rk4'
device = None

# 0.1 ... 3.95 in steps of 0.05 without 1.05, 2.05, 3.05 (BCB_latest.py:242-245)
PRODUCTION_BETA = [round(0.05 * k, 2) for k in range(2, 80) if k not in (21, 41, 61)]

_dropin.install(globals(), "BCBL")


def production_sweep(step_functions=("lin", "sqrt", "cbrt", "cerf_3"),
                     dims=tuple(np.arange(3, 71 + 1, 2))):
    """The __main__ loop of BCB_latest.py:309-320."""
    global step_function, dimension
    for step in step_functions:
        step_function = step
        for dim in dims:
            dimension = int(dim)
            parallel_routine()                             # noqa: F821 (installed above)
