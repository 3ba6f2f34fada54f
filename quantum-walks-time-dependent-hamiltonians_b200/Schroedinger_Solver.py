"""Drop-in for the hot path of the reference's ``Schroedinger_Solver.py`` (cycle graph).

Same function names, argument meaning, return shapes and module globals
(Schroedinger_Solver.py:16-18): set ``dimension``, ``step_function`` (1 = lin, 2 = sqrt,
3 = cbrt) and call ``grid_probability_evaluation`` / ``evaluate_probability`` /
``solve_schrodinger_equation`` / ``single_evaluation_benchmark`` / ``optimization``.  The
adiabatic-theorem diagnostics (``compute_energy_diff``, ``compute_gamma``,
``adiabatic_theorem_check``; Schroedinger_Solver.py:146-280) run on the device too, from the
secular equation of the rank-one perturbation instead of dense eigensolves;
``reference_derivative = True`` keeps the reference's cube-root slope 1/(3 cbrt(s)).  The plots
of that script are outside the accelerated path.
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
reference_derivative = True  # dH/ds of the cube-root schedule as the reference codes it
device = None
mirror = True                # symmetry-reduced state (QW_FLAG_MIRROR), exact for the uniform start every
                             # caller uses (BCB.py:179-180): half the ring / (psi_w, psi_other) on the
                             # complete graph, same p, norm and psi; False integrates every site

_dropin.install(globals(), "SS")
