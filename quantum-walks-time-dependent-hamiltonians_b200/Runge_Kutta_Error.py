"""Drop-in for the reference's ``Runge_Kutta_Error.py``: identical to
``Schroedinger_Solver`` except that ``solve_schrodinger_equation`` returns the norm
<psi|psi> (Runge_Kutta_Error.py:103), the integrator-accuracy diagnostic of that script."""

from . import _dropin

dimension = None
step_function = 1
rtolerance = 1e-6
atolerance = 1e-6
n_steps = None
step_size = 1e-3
gamma_on = "oracle"
integrator = "rk4"           # 'rk45': SciPy's adaptive RK45 as the reference ships it
device = None
mirror = True                # symmetry-reduced state (QW_FLAG_MIRROR), exact for the uniform start every
                             # caller uses (BCB.py:179-180): half the ring / (psi_w, psi_other) on the
                             # complete graph, same p, norm and psi; False integrates every site

_dropin.install(globals(), "RKE")


def norm_defect(time, beta):
    """1 - <psi|psi>, the 'RK45 Error' line of Runge_Kutta_Error.py:481-483."""
    return 1 - solve_schrodinger_equation(time, beta)     # noqa: F821 (installed above)
