"""numpy restatement of the fixed-step RK4 path.  TEST INFRASTRUCTURE ONLY.

Step-rule semantics (fixed from QW_Search.py:149-164):

* ``n_steps=S``    h = T/S, t_n = n*h (multiplication, not accumulation), stage times
                   t_n, t_n+h/2, t_n+h/2, t_n+h, tau = t/T by division (BCB.py:76).
* ``step_size=h``  stub-faithful: n = int(T/h) steps of size h, ending at n*h <= T.
* T == 0 (or zero steps) leaves psi = psi0, so p = 1/N (BCB.py:72-73).

Output p is renormalised, p = |psi_w|^2 / ||psi||^2 (BCB.py:189-199); ||psi||^2 is
returned next to it (Runge_Kutta_Error.py:103).
"""

import numpy as np

from .defs import (TOPOLOGY_COMPLETE, TOPOLOGY_CYCLE, SCHEDULES, GAMMA_ON_ORACLE,
                   GAMMA_ON_LAPLACIAN, schedule_id, topology_id, gamma_on_id,
                   default_target)

_STATIC = SCHEDULES["static"]


def schedule_g(tau, schedule):
    """g_T(tau) -- BCB.py:75-87, QW_Search.py:104-109, thesis Eq. 2.5-2.6."""
    s = schedule_id(schedule)
    tau = np.asarray(tau, dtype=np.float64)
    if s == SCHEDULES["lin"]:
        return tau.copy()
    if s == SCHEDULES["sqrt"]:
        return np.sqrt(tau)
    if s == SCHEDULES["cbrt"]:
        return np.cbrt(tau)
    if s == SCHEDULES["cerf_3"]:
        return 0.5 * (1 + (2 * tau - 1) ** 3)
    if s == SCHEDULES["cerf_5"]:
        return 0.5 * (1 + (2 * tau - 1) ** 5)
    if s == SCHEDULES["cerf_7"]:
        return 0.5 * (1 + (2 * tau - 1) ** 7)
    if s == _STATIC:
        return np.ones_like(tau)
    raise ValueError("schedule value not defined: %r" % (schedule,))


def operator_scalars(tau, gamma, schedule, gamma_on):
    """H(t) = a(t) L + d(t) |w><w|.  Returns (a, d).

    gamma_on='oracle'    (BCB.py:93-97):  a = 1-g,        d = -g*gamma
    gamma_on='laplacian' (thesis Eq 2.14): a = (1-g)*gamma, d = -g
    static (BCB.py:101-135):  H = L - gamma|w><w|  (resp. gamma*L - |w><w|).
    """
    s = schedule_id(schedule)
    go = gamma_on_id(gamma_on)
    gamma = np.asarray(gamma, dtype=np.float64)
    if s == _STATIC:
        one = np.ones_like(np.asarray(tau, dtype=np.float64) * gamma)
        if go == GAMMA_ON_ORACLE:
            return one, -gamma * one
        return gamma * one, -one
    g = schedule_g(tau, s)
    if go == GAMMA_ON_ORACLE:
        return (1 - g), -g * gamma
    return (1 - g) * gamma, -g * np.ones_like(gamma)


def laplacian_dense(n, topology):
    """Dense L, restating BCB.py:30-64 (cycle: N >= 3 only)."""
    topo = topology_id(topology)
    if topo == TOPOLOGY_COMPLETE:
        return n * np.eye(n) - np.ones((n, n))
    if n < 3:
        raise ValueError("cycle graph needs N >= 3")
    lap = 2.0 * np.eye(n)
    idx = np.arange(n)
    lap[idx, (idx + 1) % n] -= 1
    lap[idx, (idx - 1) % n] -= 1
    return lap


def step_plan(T, n_steps=None, step_size=None):
    """(S, h) for one point under either step rule."""
    T = float(T)
    if (n_steps is None) == (step_size is None):
        raise ValueError("give exactly one of n_steps / step_size")
    if step_size is not None:
        h = float(step_size)
        return int(T / h), h
    S = int(n_steps)
    if T == 0.0 or S == 0:
        return 0, 0.0
    return S, T / S


def _finish(psi, w):
    norm = float(np.sum(psi.real ** 2 + psi.imag ** 2))
    pw = float(psi[w].real ** 2 + psi[w].imag ** 2)
    return pw / norm, norm


def _rk4_loop(rhs, y, T, S, h):
    """Classic RK4 (weights 1/6,1/3,1/3,1/6) -- the loop QW_Search.py:149-164 sketches."""
    for n in range(S):
        t = n * h
        k1 = rhs(t, y)
        k2 = rhs(t + 0.5 * h, y + (0.5 * h) * k1)
        k3 = rhs(t + 0.5 * h, y + (0.5 * h) * k2)
        k4 = rhs(t + h, y + h * k3)
        y = y + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    return y


def rk4_L0(ref_module, T, beta, n_steps=None, step_size=None):
    """RK4 over the reference's *own* RHS ``schrodinger_equation(t,y,beta,T)``.

    ``ref_module`` comes from ``ref_loader.load('BCB', dimension=..., ...)``; H(t), the
    schedules, the oracle site, the -i factor and the initial state are all the
    reference's code (BCB.py:28-99, 163-174, 179-180).  Returns (p, norm, psi).
    """
    n = int(ref_module.dimension)
    S, h = step_plan(T, n_steps, step_size)
    y = np.empty(n, dtype=complex)
    y.fill(1 / np.sqrt(n))

    def rhs(t, yy):
        return np.asarray(ref_module.schrodinger_equation(t, yy, beta, T), dtype=complex)

    y = _rk4_loop(rhs, y, T, S, h)
    p, norm = _finish(y, default_target(n))
    return p, norm, y


def rk4_L0_static(ref_module, time, gamma, n_steps=None, step_size=None):
    """RK4 over the reference's dense time-independent H (BCB.py:101-135)."""
    n = int(ref_module.dimension)
    S, h = step_plan(time, n_steps, step_size)
    H = ref_module.generate_time_independent_hamiltonian(n, gamma)
    y = np.empty(n, dtype=complex)
    y.fill(1 / np.sqrt(n))
    y = _rk4_loop(lambda t, yy: -1j * (H @ yy), y, time, S, h)
    p, norm = _finish(y, default_target(n))
    return p, norm, y


def rk4_L1(n, topology, schedule, T, gamma, n_steps=None, step_size=None,
           gamma_on=GAMMA_ON_ORACLE, target=None):
    """Dense restatement: H(t) = a L + d|w><w| built like BCB.py:93-97, then H @ y."""
    w = default_target(n) if target is None else int(target)
    lap = laplacian_dense(n, topology)
    S, h = step_plan(T, n_steps, step_size)
    y = np.full(n, 1 / np.sqrt(n), dtype=complex)

    def rhs(t, yy):
        a, d = operator_scalars(t / T, gamma, schedule, gamma_on)
        H = float(a) * lap
        H[w, w] += float(d)
        return -1j * (H @ yy)

    y = _rk4_loop(rhs, y, T, S, h)
    p, norm = _finish(y, w)
    return p, norm, y


def apply_rhs(u, a, d, topology, w):
    """-i (a L + d|w><w|) u for a batch u[P,N]; a, d broadcast as [P,1]."""
    topo = topology_id(topology)
    n = u.shape[-1]
    if topo == TOPOLOGY_CYCLE:
        lu = 2 * u - np.roll(u, 1, axis=-1) - np.roll(u, -1, axis=-1)
    else:
        lu = n * u - u.sum(axis=-1, keepdims=True)
    v = a * lu
    v[..., w] += (d * u[..., w:w + 1])[..., 0]
    return -1j * v


def rk4_L2(n, topology, schedule, T, gamma, n_steps=None, step_size=None,
           gamma_on=GAMMA_ON_ORACLE, target=None, return_psi=False):
    """Implicit-operator restatement, batched over points.

    ``T`` and ``gamma`` are equal-length 1-D arrays (one entry per point); ``n_steps`` is a
    scalar or per-point array.  Returns (p[P], norm[P]) and optionally psi[P,N].
    """
    T = np.atleast_1d(np.asarray(T, dtype=np.float64))
    gamma = np.atleast_1d(np.asarray(gamma, dtype=np.float64))
    T, gamma = np.broadcast_arrays(T, gamma)
    P = T.shape[0]
    w = default_target(n) if target is None else int(target)
    topo = topology_id(topology)
    if topo == TOPOLOGY_CYCLE and n < 3:
        raise ValueError("cycle graph needs N >= 3")

    if step_size is not None:
        h = np.full(P, float(step_size))
        S = np.array([int(t / float(step_size)) for t in T], dtype=np.int64)
    else:
        S = np.broadcast_to(np.asarray(n_steps, dtype=np.int64), (P,)).copy()
        S[T == 0.0] = 0
        with np.errstate(divide="ignore", invalid="ignore"):
            h = np.where(S > 0, T / np.maximum(S, 1), 0.0)

    y = np.full((P, n), 1 / np.sqrt(n), dtype=complex)
    hc = h[:, None]
    Tc = np.where(T == 0.0, 1.0, T)[:, None]
    gc = gamma[:, None]

    def coef(t):
        return operator_scalars(t / Tc, gc, schedule, gamma_on)

    for nstep in range(int(S.max()) if P else 0):
        live = (S > nstep)[:, None]
        t = nstep * hc
        a, d = coef(t)
        k1 = apply_rhs(y, a, d, topo, w)
        a, d = coef(t + 0.5 * hc)
        k2 = apply_rhs(y + (0.5 * hc) * k1, a, d, topo, w)
        k3 = apply_rhs(y + (0.5 * hc) * k2, a, d, topo, w)
        a, d = coef(t + hc)
        k4 = apply_rhs(y + hc * k3, a, d, topo, w)
        ynew = y + (hc / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
        y = np.where(live, ynew, y)

    norm = np.sum(y.real ** 2 + y.imag ** 2, axis=1)
    pw = y[:, w].real ** 2 + y[:, w].imag ** 2
    p = pw / norm
    if return_psi:
        return p, norm, y
    return p, norm


def heatmap_L2(n, topology, schedule, time_array, beta_array, **kw):
    """p[n_beta, n_T] in the reference's layout (BCB.py:239-259, SS:426-444)."""
    time_array = np.asarray(time_array, dtype=np.float64)
    beta_array = np.asarray(beta_array, dtype=np.float64)
    bb, tt = np.meshgrid(beta_array, time_array, indexing="ij")
    p, norm = rk4_L2(n, topology, schedule, tt.ravel(), bb.ravel(), **kw)
    return p.reshape(bb.shape), norm.reshape(bb.shape)


def complete_two_variable(n, schedule, T, gamma, n_steps, gamma_on=GAMMA_ON_ORACLE):
    """Symmetric-subspace check for the complete graph (invariant i).

    All non-target amplitudes stay equal, so with x = psi_w and z = any other amplitude
      (L psi)_w = (N-1)(x - z),  (L psi)_other = z - x
    and the N-site RK4 must agree with RK4 on (x, z) to rounding.  Returns (p, norm).
    """
    S, h = step_plan(T, n_steps=n_steps)
    y = np.array([1 / np.sqrt(n), 1 / np.sqrt(n)], dtype=complex)

    def rhs(t, yy):
        a, d = operator_scalars(t / T, gamma, schedule, gamma_on)
        x, z = yy
        return -1j * np.array([float(a) * (n - 1) * (x - z) + float(d) * x,
                               float(a) * (z - x)])

    y = _rk4_loop(rhs, y, T, S, h)
    norm = abs(y[0]) ** 2 + (n - 1) * abs(y[1]) ** 2
    return abs(y[0]) ** 2 / norm, norm


def robustness(probability, beta_index, time_index, variation):
    """Robustness.py:25-31 restated: mean drop at +-variation grid cells, 4 decimals."""
    p0 = probability[beta_index, time_index]
    r_pos = p0 - probability[beta_index + variation, time_index]
    r_neg = p0 - probability[beta_index - variation, time_index]
    return round(float((r_pos + r_neg) / 2), 4)


def first_argmax(values):
    """Robustness.py:41-45: strict '>' scan from index 0 against 0 -> first maximum."""
    best, idx = 0.0, None
    for j, v in enumerate(values):
        if v > best:
            best, idx = v, j
    return idx


def ti_exact(n, time, gamma, gamma_on=GAMMA_ON_ORACLE, target=None, topology="circular"):
    """exp(-i H t)|s> by dense diagonalisation: the reference's time-independent comparator
    (BCB.py:137-159 uses scipy.linalg.expm on the same H = L - gamma|w><w|).
    Returns (|psi_w|^2, ||psi||^2, psi)."""
    w = default_target(n) if target is None else int(target)
    lap = laplacian_dense(n, topology)
    if gamma_on_id(gamma_on) == GAMMA_ON_ORACLE:
        H = lap.copy()
        H[w, w] += -gamma
    else:
        H = gamma * lap
        H[w, w] += -1.0
    lam, V = np.linalg.eigh(H)
    psi0 = np.full(n, 1 / np.sqrt(n))
    psi = V @ (np.exp(-1j * lam * time) * (V.T @ psi0))
    return float(abs(psi[w]) ** 2), float(np.vdot(psi, psi).real), psi


def rk45_scipy(n, topology, schedule, T, gamma, rtol=1e-6, atol=1e-6, gamma_on=GAMMA_ON_ORACLE,
               target=None):
    """The reference's shipped integrator restated: scipy solve_ivp RK45 (BCB.py:177-190) over
    the implicit right-hand side.  Returns (p, norm, psi, nfev)."""
    from scipy.integrate import solve_ivp
    w = default_target(n) if target is None else int(target)
    topo = topology_id(topology)
    y0 = np.full(n, 1 / np.sqrt(n), dtype=complex)
    if T == 0:
        p, norm = _finish(y0, w)
        return p, norm, y0, 0

    def fun(t, y):
        if t == 0 and schedule_id(schedule) != _STATIC:          # BCB.py:72-73
            a, d = (1.0 if gamma_on_id(gamma_on) == GAMMA_ON_ORACLE else gamma), 0.0
        else:
            a, d = operator_scalars(t / T, gamma, schedule, gamma_on)
        return apply_rhs(y[None, :], float(a), float(d), topo, w)[0]

    sol = solve_ivp(fun, [0.0, T], y0, method="RK45", atol=atol, rtol=rtol)
    psi = sol.y[:, -1]
    p, norm = _finish(psi, w)
    return p, norm, psi, sol.nfev


def spectrum_dense(n, topology, schedule, s, gamma, gamma_on=GAMMA_ON_ORACLE, target=None):
    """Sorted eigenvalues of H(s) = a L + d |w><w| by dense diagonalisation, restating
    compute_eigenvalues of Eigenvalues_Crossing.py:59-71 (time_dependent_hamiltonian, lines
    58-63, + scipy.linalg.eig + np.sort; H is real symmetric, so eigvalsh returns the same
    numbers).  The reference's file is Python 2 (``hamiltonian[(dimension-1)/2]``) and cannot be
    executed here: parity for this row is pinned to this restatement only."""
    w = (n - 1) // 2 if target is None else target
    a, d = (float(x) for x in operator_scalars(s, gamma, schedule, gamma_on))
    H = a * laplacian_dense(n, topology)
    H[w, w] += d
    return np.linalg.eigvalsh(H)
