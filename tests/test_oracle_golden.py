"""CPU: the oracle restatements against fixtures produced by the reference's own code
(tests/golden/make_golden.py) and against each other (L0 == L1 == L2 == C)."""

import numpy as np
import pytest

from oracle import c_oracle, ref_loader, rk4_numpy as rk
from oracle.defs import default_target

PSI_TOL = 5e-14     # restatement vs reference-driven RK4: rounding only
P_TOL = 5e-14


def _rule(case):
    return {k: v for k, v in case["rule"].items()}


def test_dense_hamiltonian_matches_reference(golden):
    meta, arr = golden
    for c in meta["hamiltonians"]:
        H_ref = arr[c["key"]]
        n, w = c["n"], default_target(c["n"])
        lap = rk.laplacian_dense(n, c["topology"])
        if c["t"] == 0 or c["T"] == 0:           # BCB.py:72-73 guard: H = L
            H = lap
        else:
            a, d = rk.operator_scalars(c["t"] / c["T"], c["beta"], c["schedule"], "oracle")
            H = float(a) * lap
            H[w, w] += float(d)
        assert np.abs(H - H_ref).max() <= 1e-15 * max(1.0, np.abs(H_ref).max()), c


def test_rhs_matches_reference(golden):
    meta, arr = golden
    for c in meta["rhs"]:
        y, dy_ref = arr["rhs_y_%d" % c["key"]], arr["rhs_dy_%d" % c["key"]]
        a, d = rk.operator_scalars(c["t"] / c["T"], c["beta"], c["schedule"], "oracle")
        dy = rk.apply_rhs(y[None, :], float(a), float(d), c["topology"],
                          default_target(c["n"]))[0]
        assert np.abs(dy - dy_ref).max() <= 2e-13 * np.abs(dy_ref).max(), c


def test_rk4_restatements_match_L0(golden):
    meta, arr = golden
    for c in meta["rk4_L0"]:
        psi_ref = arr["L0_psi_%d" % c["key"]]
        args = (c["n"], c["topology"], c["schedule"])
        p1, n1, psi1 = rk.rk4_L1(*args, c["T"], c["beta"], **_rule(c))
        p2, n2, psi2 = rk.rk4_L2(*args, [c["T"]], [c["beta"]], return_psi=True, **_rule(c))
        pc, nc, psic, _ = c_oracle.rk4_batch(*args, [c["T"]], [c["beta"]], return_psi=True,
                                             **_rule(c))
        for p, nrm, psi in ((p1, n1, psi1), (p2[0], n2[0], psi2[0]), (pc[0], nc[0], psic[0])):
            assert abs(p - c["p"]) <= P_TOL, c
            assert abs(nrm - c["norm"]) <= P_TOL, c
            assert np.abs(psi - psi_ref).max() <= PSI_TOL, c


def test_rk45_reference_is_close(golden):
    """The reference as shipped (RK45, tol 1e-6) vs converged RK4: loose, informational
    bound (SURVEY.md finding 2: ~3e-6 for lin/cerf, up to ~1e-4 for sqrt/cbrt)."""
    meta, _ = golden
    for c in meta["rk45"]:
        p, _ = rk.rk4_L2(c["n"], c["topology"], c["schedule"], [c["T"]], [c["beta"]],
                         n_steps=8192)
        tol = 2e-4 if c["schedule"] in ("sqrt", "cbrt") else 3e-5
        assert abs(p[0] - c["p"]) <= tol, (c, p[0])


def test_static_comparator(golden):
    meta, _ = golden
    for c in meta["static"]:
        p, nrm, _ = c_oracle.rk4_batch(c["n"], "circular", "static", [c["time"]],
                                       [c["gamma"]], n_steps=c["n_steps"])
        assert abs(p[0] - c["p_rk4"]) <= P_TOL, c
        assert abs(p[0] - c["p_expm"]) <= 1e-8, c       # RK4 truncation vs exact expm


def test_heatmap_layout_beta_rows_time_cols(golden):
    meta, arr = golden
    for c in meta["heatmaps"]:
        t = np.linspace(*c["time"][:2], int(c["time"][2]))
        b = np.linspace(*c["beta"][:2], int(c["beta"][2]))
        if c["kind"] == "L0":
            p, _ = rk.heatmap_L2(c["n"], c["topology"], c["schedule"], t, b,
                                 n_steps=c["n_steps"])
            pc, _ = c_oracle.heatmap(c["n"], c["topology"], c["schedule"], t, b,
                                     n_steps=c["n_steps"])
            assert np.abs(p - arr[c["key"]]).max() <= P_TOL
            assert np.abs(pc - arr[c["key"]]).max() <= P_TOL
        else:
            p, _ = c_oracle.heatmap(c["n"], c["topology"], c["schedule"], t, b, n_steps=4096)
            assert p.shape == arr[c["key"]].shape == (len(b), len(t))
            assert np.abs(p - arr[c["key"]]).max() <= 3e-5


def test_robustness_formula(golden):
    meta, arr = golden
    for c in meta["robustness"]:
        assert rk.robustness(arr["rob_prob"], c["beta_index"], c["time_index"],
                             c["variation"]) == c["R"]


def test_known_answers():
    # p(T=0) = 1/N (BCB.py:72 guard)
    p, nrm = rk.rk4_L2(16, "circular", "lin", [0.0], [1.0], n_steps=32)
    assert p[0] == pytest.approx(1 / 16, abs=1e-15) and nrm[0] == pytest.approx(1, abs=1e-15)
    # analytic Grover point, thesis Eq. 1.28 in the code's placement: H = L - N|w><w|,
    # T = pi/(2 sqrt N) -> p = 1
    n = 51
    p, nrm, _ = c_oracle.rk4_batch(n, "complete", "static", [np.pi / (2 * np.sqrt(n))],
                                   [float(n)], n_steps=2000)
    assert abs(p[0] - 1.0) < 1e-12
    # ... and with gamma on the Laplacian (thesis Eq. 2.14): H = (1/N) L - |w><w|,
    # T = pi sqrt(N) / 2
    p, nrm, _ = c_oracle.rk4_batch(n, "complete", "static", [np.pi * np.sqrt(n) / 2],
                                   [1.0 / n], n_steps=4000, gamma_on="laplacian")
    assert abs(p[0] - 1.0) < 1e-10


def test_complete_graph_two_variable_invariant():
    for go in ("oracle", "laplacian"):
        n = 4096
        gamma, T, S = (0.7, 0.05, 400) if go == "oracle" else (1.2 / n, 60.0, 600)
        p, nrm, psi, _ = c_oracle.rk4_batch(n, "complete", "cerf_3", [T], [gamma],
                                            n_steps=S, gamma_on=go, return_psi=True)
        p2, n2 = rk.complete_two_variable(n, "cerf_3", T, gamma, S, gamma_on=go)
        assert abs(p[0] - p2) <= 1e-12 and abs(nrm[0] - n2) <= 1e-11
        others = np.delete(psi[0], default_target(n))
        assert np.abs(others - others[0]).max() <= 1e-15


def test_cycle_mirror_symmetry():
    for n in (64, 51):
        _, _, psi, _ = c_oracle.rk4_batch(n, "circular", "lin", [30.0], [1.1], n_steps=512,
                                          return_psi=True)
        w = default_target(n)
        k = np.arange(1, n // 2)
        assert np.abs(psi[0][(w + k) % n] - psi[0][(w - k) % n]).max() <= 1e-14


def test_step_size_rule_and_ragged_batch():
    T = np.array([0.0, 0.004, 1.0, 2.5001, 7.0])
    g = np.array([1.0, 0.5, 2.0, 0.1, 1.5])
    p2, n2 = rk.rk4_L2(9, "circular", "sqrt", T, g, step_size=1e-2)
    pc, nc, _ = c_oracle.rk4_batch(9, "circular", "sqrt", T, g, step_size=1e-2)
    assert np.abs(p2 - pc).max() <= P_TOL and np.abs(n2 - nc).max() <= P_TOL
    assert p2[0] == pytest.approx(1 / 9, abs=1e-15) and p2[1] == pytest.approx(1 / 9, abs=1e-15)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree absent")
def test_live_reference_rhs_and_L0():
    """Only in the build container: drive the live reference, not just its fixtures."""
    bcb = ref_loader.load("BCB", dimension=12, step_function="cerf_3", topology="circular")
    p0, n0, psi0 = rk.rk4_L0(bcb, 6.0, 1.7, n_steps=200)
    p2, n2, psi2 = rk.rk4_L2(12, "circular", "cerf_3", [6.0], [1.7], n_steps=200,
                             return_psi=True)
    assert abs(p0 - p2[0]) <= P_TOL and np.abs(psi0 - psi2[0]).max() <= PSI_TOL


def test_ti_exact_oracle_matches_reference_expm(golden):
    meta, _ = golden
    for c in meta["static"]:
        p, nrm, _ = rk.ti_exact(c["n"], c["time"], c["gamma"])
        assert abs(p - c["p_expm"]) <= 1e-13 and abs(nrm - 1) <= 1e-13, c


def test_rk45_restatement_matches_the_shipped_reference(golden):
    """SciPy RK45 over the stencil RHS vs the reference's own evaluate_probability output."""
    meta, arr = golden
    for c in meta["rk45"]:
        p, _, _, _ = rk.rk45_scipy(c["n"], c["topology"], c["schedule"], c["T"], c["beta"])
        assert abs(p - c["p"]) <= 1e-10, (c, p)
