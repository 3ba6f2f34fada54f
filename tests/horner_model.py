"""numpy model of the nested (Horner) form of one RK4 step that the `*_horner` CUDA kernels
use.  TEST INFRASTRUCTURE ONLY (it is checked against oracle/rk4_numpy.py, which follows
the reference: BCB.py:163-190, QW_Search.py:149-164).

With F_s = a_s M + d_s Q, M = -iL, Q = -i|w><w| the RK4 step of y' = F(t) y is

    y+ = y + sum_s b_s F_s u_s,   u_1 = y,  u_{s+1} = y + c_s F_s u_s.

Without the oracle term every F_s is a multiple of M, the step operator is the polynomial
I + b1 M + b2 M^2 + b3 M^3 + b4 M^4 and can be evaluated as four nested axpy's

    z4 = y + r4 M y,  z3 = y + r3 M z4,  z2 = y + r2 M z3,  y+ = y + r1 M z2,

r1 = b1, r2 = b2/b1, r3 = b3/b2, r4 = b4/b3: 24 instead of 30 FP64 instructions per
site-update and no accumulator vector.  The oracle term does not commute with L, but every
word of the step operator that contains a Q is a multiple of M^k |w>, k = 0..3, so the
difference is restored *exactly* by adding a complex scalar s_l to the marked vertex of z4,
z3, z2 and y+ (the nested M's spread it).  The four scalars are linear in
(y_w, (Ly)_w, (L^2 y)_w, (L^3 y)_w) with coefficients that depend on the step's scalars
only, so they are tabulated per step (`step_coefficients`) next to r1..r4.
"""

import numpy as np

from oracle.defs import TOPOLOGY_CYCLE, topology_id
from oracle.rk4_numpy import operator_scalars


def _div(x, y):
    return x / y if y != 0.0 else 0.0 * x


def step_coefficients(h, a1, a2, a4, d1, d2, d4, lww, l2ww):
    """rho[4] (r1..r4) and C[l] (l = 1..4): coefficient vectors over the basis
    (y_w, (Ly)_w, (L^2y)_w, (L^3y)_w) of the scalar injected after level l
    (l = 4 innermost).  lww = L_ww, l2ww = (L^2)_ww."""
    A1, A2, A3 = 0.5 * h * a1, 0.5 * h * a2, h * a2
    B1, B2, B3, B4 = h / 6 * a1, h / 3 * a2, h / 3 * a2, h / 6 * a4
    DA1, DA2, DA3 = 0.5 * h * d1, 0.5 * h * d2, h * d2
    DB1, DB2, DB3, DB4 = h / 6 * d1, h / 3 * d2, h / 3 * d2, h / 6 * d4
    b1 = B1 + B2 + B3 + B4
    b2 = B2 * A1 + B3 * A2 + B4 * A3
    b3 = B3 * A2 * A1 + B4 * A3 * A2
    b4 = B4 * A3 * A2 * A1
    rho = np.array([b1, _div(b2, b1), _div(b3, b2), _div(b4, b3)])

    e = np.eye(4, dtype=complex)
    mww, m2ww = -1j * lww, -l2ww                    # (M)_ww, (M^2)_ww
    q1 = e[0]
    al1 = -1j * DA1 * q1
    q2 = e[0] - 1j * A1 * e[1] + al1
    al2 = -1j * DA2 * q2
    q3 = e[0] - 1j * A2 * e[1] - A2 * A1 * e[2] + A2 * mww * al1 + al2
    al3 = -1j * DA3 * q3
    q4 = (e[0] - 1j * A3 * e[1] - A3 * A2 * e[2] + 1j * A3 * A2 * A1 * e[3]
          + A3 * A2 * m2ww * al1 + A3 * mww * al2 + al3)
    k0 = -1j * (DB1 * q1 + DB2 * q2 + DB3 * q3 + DB4 * q4)
    k1 = B2 * al1 + B3 * al2 + B4 * al3
    k2 = B3 * A2 * al1 + B4 * A3 * al2
    k3 = B4 * A3 * A2 * al1
    C = {1: k0, 2: _div(k1, b1), 3: _div(k2, b2), 4: _div(k3, b3)}
    return rho, C


def lap(u, topo):
    if topo == TOPOLOGY_CYCLE:
        return 2 * u - np.roll(u, 1) - np.roll(u, -1)
    return u.shape[0] * u - u.sum()


def horner_step(y, w, topo, rho, C):
    n = y.shape[0]
    m = [y[w]]
    v = y
    for _ in range(3):
        v = lap(v, topo)
        m.append(v[w])
    m = np.array(m)
    z = y
    for level in (4, 3, 2, 1):
        z = y - 1j * rho[level - 1] * lap(z, topo)
        z[w] += C[level] @ m
    return z


def rk4_horner(n, topology, schedule, T, gamma, n_steps, gamma_on, target=None):
    topo = topology_id(topology)
    w = (n - 1) // 2 if target is None else target
    y = np.full(n, 1 / np.sqrt(n), dtype=complex)
    if T == 0 or n_steps <= 0:
        return 1.0 / n, 1.0, y
    h = T / n_steps
    lww = 2.0 if topo == TOPOLOGY_CYCLE else n - 1.0
    l2ww = 6.0 if topo == TOPOLOGY_CYCLE else n * (n - 1.0)
    for s in range(n_steps):
        t = s * h
        a1, d1 = (float(x) for x in operator_scalars(t / T, gamma, schedule, gamma_on))
        a2, d2 = (float(x) for x in operator_scalars((t + 0.5 * h) / T, gamma, schedule, gamma_on))
        a4, d4 = (float(x) for x in operator_scalars((t + h) / T, gamma, schedule, gamma_on))
        rho, C = step_coefficients(h, a1, a2, a4, d1, d2, d4, lww, l2ww)
        y = horner_step(y, w, topo, rho, C)
    norm = float(np.sum(np.abs(y) ** 2))
    return float(np.abs(y[w]) ** 2) / norm, norm, y


# ---- cycle graph: the run-time form of the correction (csrc/qw_horner_level.cuh, HornerStep) ----

def cycle_window_table(C):
    """Per level l the four complex coefficients of (y_w, y_{w-1}, y_{w-2}, y_{w-3}) in
    t_l = y_w + sigma_l, as horner_step_store tabulates them: change of basis to the neighbour
    sums s_q = y_{w-q} + y_{w+q} (Ly = 2y - s1, L^2y = 6y - 4s1 + s2, L^3y = 20y - 15s1 + 6s2 - s3),
    the identity folded into the coefficient of y_w, and s_q = 2 y_{w-q} (the state is mirror
    symmetric about the marked vertex) folded into the others."""
    basis = np.array([[1, 0, 0, 0], [2, -1, 0, 0], [6, -4, 1, 0], [20, -15, 6, -1]], dtype=complex)
    tab = {}
    for level in (1, 2, 3, 4):
        c = np.asarray(C[level], dtype=complex) @ basis
        c[0] += 1.0
        c[1:] *= 2.0
        tab[level] = c
    return tab


def horner_step_addend(y, w, rho, tab):
    """One step as the kernels run it: the level's axpy at the marked vertex takes t_l (a dot
    product over the one-sided window y_w .. y_{w-3} of the step's start) as its addend."""
    n = y.shape[0]
    win = np.array([y[w], y[(w - 1) % n], y[(w - 2) % n], y[(w - 3) % n]])
    z = y
    for level in (4, 3, 2, 1):
        mz = -1j * rho[level - 1] * lap(z, TOPOLOGY_CYCLE)
        znew = y + mz
        znew[w] = tab[level] @ win + mz[w]
        z = znew
    return z


def rk4_horner_addend(n, schedule, T, gamma, n_steps, gamma_on, target=None):
    w = (n - 1) // 2 if target is None else target
    y = np.full(n, 1 / np.sqrt(n), dtype=complex)
    h = T / n_steps
    asym = 0.0
    for s in range(n_steps):
        t = s * h
        a1, d1 = (float(x) for x in operator_scalars(t / T, gamma, schedule, gamma_on))
        a2, d2 = (float(x) for x in operator_scalars((t + 0.5 * h) / T, gamma, schedule, gamma_on))
        a4, d4 = (float(x) for x in operator_scalars((t + h) / T, gamma, schedule, gamma_on))
        rho, C = step_coefficients(h, a1, a2, a4, d1, d2, d4, 2.0, 6.0)
        y = horner_step_addend(y, w, rho, cycle_window_table(C))
        k = np.arange(1, n // 2 + 1)
        asym = max(asym, float(np.abs(y[(w + k) % n] - y[(w - k) % n]).max()))
    norm = float(np.sum(np.abs(y) ** 2))
    return float(np.abs(y[w]) ** 2) / norm, norm, y, asym


# ---- complete graph: the marked-vertex subsystem as a linear map (csrc/qw_horner.cu, CSMap) ----

def complete_subsystem_map(rho, C, n):
    """Ten outputs of one step as out = qa * y_w + qb * S(y): the four level means
    (S(y), S(z4), S(z3), S(z2)) / n, the corrections sigma_4..sigma_1, then y_w and S(y) after the
    step.  Obtained as in the kernel: the two basis vectors go through the operations the
    vector code applies to the marked vertex."""
    def push(yw, sy):
        m1 = n * yw - sy                                         # (L y)_w
        g = {l: C[l][0] * yw + (C[l][1] + n * C[l][2] + n * n * C[l][3]) * m1
             for l in (1, 2, 3, 4)}                              # L^2y_w = n Ly_w, L^3y_w = n^2 Ly_w
        out = [sy / n]
        v = m1
        for l in (4, 3, 2):
            s_l = sy + g[l]
            out.append(s_l / n)
            z = yw - 1j * rho[l - 1] * v + g[l]
            v = n * z - s_l
        out += [g[4], g[3], g[2], g[1]]
        out += [yw - 1j * rho[0] * v + g[1], sy + g[1]]
        return np.array(out)
    return push(1.0 + 0j, 0j), push(0j, 1.0 + 0j)


def rk4_complete_mapped(n, schedule, T, gamma, n_steps, gamma_on, target=None):
    """The complete-graph kernels' scheme: the vector code never sums anything; x = (y_w, S(y)) is
    advanced by the tabulated map, the levels subtract the level mean (y - i rho n (z - S/n)) and
    the marked vertex receives sigma_l.  S(y) is never re-anchored."""
    w = (n - 1) // 2 if target is None else target
    y = np.full(n, 1 / np.sqrt(n), dtype=complex)
    x = np.array([y[w], y.sum()])
    h = T / n_steps
    for s in range(n_steps):
        t = s * h
        a1, d1 = (float(v) for v in operator_scalars(t / T, gamma, schedule, gamma_on))
        a2, d2 = (float(v) for v in operator_scalars((t + 0.5 * h) / T, gamma, schedule, gamma_on))
        a4, d4 = (float(v) for v in operator_scalars((t + h) / T, gamma, schedule, gamma_on))
        rho, C = step_coefficients(h, a1, a2, a4, d1, d2, d4, n - 1.0, n * (n - 1.0))
        qa, qb = complete_subsystem_map(rho, C, n)
        out = qa * x[0] + qb * x[1]
        z = y
        for idx, level in enumerate((4, 3, 2, 1)):
            z = y - 1j * (rho[level - 1] * n) * (z - out[idx])
            z[w] += out[4 + idx]
        y = z
        x = out[8:10]
    norm = float(np.sum(np.abs(y) ** 2))
    return float(np.abs(y[w]) ** 2) / norm, norm, y, x


# ---- small complete graphs: stage form with tracked stage sums (csrc/qw_packed.cu) ------------

def rk4_complete_tracked_sums(n, schedule, T, gamma, n_steps, gamma_on, target=None):
    """Classic RK4 stages with (L x)_j = n x_j - S(x), where S of every stage input follows from
    the marked vertex alone (the column sums of L vanish): S(k_s) = -i d_s x_w."""
    w = (n - 1) // 2 if target is None else target
    y = np.full(n, 1 / np.sqrt(n), dtype=complex)
    sy = n * (1 / np.sqrt(n)) + 0j
    h = T / n_steps
    cs, bs = (0.5, 0.5, 1.0), (1 / 6, 1 / 3, 1 / 3, 1 / 6)
    for s in range(n_steps):
        t = s * h
        ad = [tuple(float(v) for v in operator_scalars(tt / T, gamma, schedule, gamma_on))
              for tt in (t, t + 0.5 * h, t + 0.5 * h, t + h)]
        x, sx, acc, sa = y, sy, y.copy(), sy
        for k in range(4):
            a, d = ad[k]
            kv = -1j * a * (n * x - sx)
            kv[w] += -1j * d * x[w]
            sk = -1j * d * x[w]
            acc = acc + h * bs[k] * kv
            sa = sa + h * bs[k] * sk
            if k < 3:
                x, sx = y + h * cs[k] * kv, sy + h * cs[k] * sk
        y, sy = acc, sa
    norm = float(np.sum(np.abs(y) ** 2))
    return float(np.abs(y[w]) ** 2) / norm, norm, y, sy
