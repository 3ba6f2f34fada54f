"""Drop-in for the hot path of the reference's ``Schroedinger_Solver.py`` (cycle graph).

Same function names, argument meaning, return shapes and module globals
(Schroedinger_Solver.py:16-18): set ``dimension``, ``step_function`` (1 = lin, 2 = sqrt,
3 = cbrt) and call ``grid_probability_evaluation`` / ``evaluate_probability`` /
``solve_schrodinger_equation`` / ``single_evaluation_benchmark``.  The SciPy optimisers,
the adiabatic-theorem check and the plots of that script are outside the accelerated path.
"""

from . import _dropin

dimension = None
step_function = 1
rtolerance = 1e-6          # used by integrator = 'rk45' only
atolerance = 1e-6
n_steps = None             # h = T / n_steps when set
step_size = 1e-3           # else n = int(T / step_size) steps (QW_Search.py:153-154)
gamma_on = "oracle"
integrator = "rk4"           # 'rk45': SciPy's adaptive RK45 as the reference ships it
device = None

_dropin.install(globals(), "SS")
