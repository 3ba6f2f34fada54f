"""numpy statement of the secular-equation form of the adiabatic diagnostics that
csrc/qw_adiabatic.cu implements (Schroedinger_Solver.py:146-253 computes the same numbers
with a dense scipy.linalg.eig per point).  TEST INFRASTRUCTURE ONLY.

H(s) = a L + d |w><w|, a = 1 - g(s), d = -g(s) beta.  Eigenvectors with phi(w) != 0 solve
1 + d G(E) = 0, G(E) = (1/N) sum_k 1/(a lambda_k - E); phi(w)^2 = 1/(d^2 G2(E));
<phi1|dH/ds|phi0> = phi1(w) phi0(w) (d' - a' d / a).
"""
import numpy as np


def schedule(s, step_function, reference=True):
    """(g, g') for the reference's step_function codes 1 = lin, 2 = sqrt, 3 = cbrt."""
    if step_function == 1:
        return s, 1.0
    if step_function == 2:
        return np.sqrt(s), 0.5 / np.sqrt(s)
    if step_function == 3:
        c = np.cbrt(s)
        return c, (1.0 / (3.0 * c) if reference else 1.0 / (3.0 * c * c))
    raise ValueError(step_function)


def gap_and_coupling(n, step_function, s, beta, reference=True):
    g, dg = schedule(s, step_function, reference)
    a, d = 1.0 - g, -g * beta
    lam = a * (2.0 - 2.0 * np.cos(2.0 * np.pi * np.arange(n) / n))
    target = -1.0 / d

    def G(E):
        return np.mean(1.0 / (lam - E))

    def root(lo, hi):
        for _ in range(300):
            mid = 0.5 * (lo + hi)
            if not (lo < mid < hi):
                break
            if G(mid) < target:
                lo = mid
            else:
                hi = mid
        return 0.5 * (lo + hi)

    e0, e1 = root(d, 0.0), root(0.0, lam[1])
    p0 = 1.0 / (d * d * np.mean(1.0 / (lam - e0) ** 2))
    p1 = 1.0 / (d * d * np.mean(1.0 / (lam - e1) ** 2))
    da, dd = -dg, -dg * beta
    return e1 - e0, np.sqrt(p0 * p1) * abs(dd - da * d / a)
