"""Adiabatic diagnostics (Schroedinger_Solver.py:146-280): the golden values
were produced by the reference's own compute_energy_diff / compute_gamma /
adiabatic_theorem_check (tests/golden/make_golden_adiabatic.py)."""
import json
import os

import numpy as np
import pytest

from adiabatic_model import gap_and_coupling

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                      "golden_adiabatic_v1.json")
SCHED = {1: "lin", 2: "sqrt", 3: "cbrt"}


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN) as fh:
        return json.load(fh)


def test_secular_model_reproduces_the_reference(gold):
    """CPU: the secular-equation form against the reference's dense eigensolves."""
    for c in gold["points"]:
        gap, cpl = gap_and_coupling(c["dimension"], c["step_function"], c["s"], c["beta"])
        assert abs(gap - c["gap"]) <= 1e-11, c
        assert abs(cpl - c["gamma"]) <= 1e-11, c


@pytest.mark.gpu
def test_device_gap_and_coupling_match_the_reference(qw, gold):
    by_cfg = {}
    for c in gold["points"]:
        by_cfg.setdefault((c["dimension"], c["step_function"]), []).append(c)
    for (dim, sf), pts in by_cfg.items():
        prob = qw.Problem(n=dim, topology="circular", schedule=SCHED[sf], n_steps=1)
        s = np.array([c["s"] for c in pts])
        b = np.array([c["beta"] for c in pts])
        gap, cpl, e0 = qw.adiabatic_batch(prob, s, b, reference_derivative=True)
        assert np.abs(gap - np.array([c["gap"] for c in pts])).max() <= 1e-10, (dim, sf)
        assert np.abs(cpl - np.array([c["gamma"] for c in pts])).max() <= 1e-10, (dim, sf)
        assert np.all(e0 < 0)


@pytest.mark.gpu
def test_device_against_dense_diagonalisation(qw):
    """Both graphs, both gamma placements, all schedules, larger dimensions: against numpy's
    symmetric eigensolver on the dense H(s) and the true dH/ds."""
    rng = np.random.default_rng(3)
    for topo, n in (("circular", 16), ("circular", 255), ("circular", 1024), ("complete", 7),
                    ("complete", 300)):
        lap = (n * np.eye(n) - np.ones((n, n))) if topo == "complete" else None
        if lap is None:
            lap = 2.0 * np.eye(n)
            i = np.arange(n)
            lap[i, (i + 1) % n] = lap[i, (i - 1) % n] = -1.0
        w = (n - 1) // 2
        for sched in ("lin", "sqrt", "cbrt", "cerf_3", "cerf_5"):
            for gamma_on in ("oracle", "laplacian"):
                s = rng.uniform(0.05, 0.95, 6)
                gam = rng.uniform(0.3, 2.0, 6) * (1.0 if (topo == "circular") == (gamma_on == "oracle")
                                                  or topo == "circular" else 1.0)
                if topo == "complete":
                    gam = gam * (n if gamma_on == "oracle" else 1.0 / n)
                prob = qw.Problem(n=n, topology=topo, schedule=sched, gamma_on=gamma_on, n_steps=1)
                gap, cpl, e0 = qw.adiabatic_batch(prob, s, gam)
                for q in range(6):
                    k = {"cerf_3": 3, "cerf_5": 5}.get(sched)
                    if sched == "lin":
                        g, dg = s[q], 1.0
                    elif sched == "sqrt":
                        g, dg = np.sqrt(s[q]), 0.5 / np.sqrt(s[q])
                    elif sched == "cbrt":
                        g, dg = np.cbrt(s[q]), 1.0 / (3.0 * np.cbrt(s[q]) ** 2)
                    else:
                        g, dg = 0.5 * (1 + (2 * s[q] - 1) ** k), k * (2 * s[q] - 1) ** (k - 1)
                    gl, gd = (1.0, gam[q]) if gamma_on == "oracle" else (gam[q], 1.0)
                    H = (1 - g) * gl * lap
                    H[w, w] += -g * gd
                    dH = -dg * gl * lap
                    dH[w, w] += -dg * gd
                    ev, vec = np.linalg.eigh(H)
                    scale = max(1.0, abs(ev).max())
                    assert abs(gap[q] - (ev[1] - ev[0])) <= 1e-10 * scale, (topo, n, sched, gamma_on)
                    assert abs(e0[q] - ev[0]) <= 1e-10 * scale
                    ref = abs(vec[:, 1] @ dH @ vec[:, 0])
                    assert abs(cpl[q] - ref) <= 1e-8 * max(1.0, abs(dH).max()), (topo, n, sched, gamma_on)


@pytest.mark.gpu
def test_adiabatic_theorem_check_drop_in(qw, gold):
    """adiabatic_theorem_check([_, _, beta]) of the drop-in module against the reference's two
    shgo runs (linear schedule).  shgo stops at its own tolerance: 1e-6 relative."""
    import importlib
    SS = importlib.import_module(qw.__name__ + ".Schroedinger_Solver")
    for c in gold["checks"]:
        SS.dimension, SS.step_function = c["dimension"], c["step_function"]
        res = SS.adiabatic_theorem_check([0.0, 0.0, c["beta"]])
        assert res[3] == c["flag"]
        assert abs(res[1] - c["gamma_max"]) <= 1e-6 * c["gamma_max"], (c, res)
        assert abs(res[2] - c["energy_min"]) <= 1e-6 * c["energy_min"], (c, res)
        assert abs(res[0] - c["adiabatic_time"]) <= 5e-6 * c["adiabatic_time"], (c, res)
        # the optimum of the batched search is at least as good as the reference's
        assert res[1] >= c["gamma_max"] * (1 - 1e-9) and res[2] <= c["energy_min"] * (1 + 1e-9)
    SS.dimension, SS.step_function = 15, 1
    assert abs(SS.compute_energy_diff(0.3, 1.02212) - 0.11848000871714828) <= 1e-12
    g = SS.compute_gamma(0.3, 1.02212)
    assert g.shape == (1, 1) and abs(g[0, 0] + 0.2188982087219179) <= 1e-12
    assert SS.generate_hamiltonian_derivative(0.3, 1.0, 1).shape == (15, 15)
