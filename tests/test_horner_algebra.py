"""CPU check of the nested (Horner) form of the RK4 step with tabulated oracle corrections
(tests/horner_model.py, the numpy statement of csrc/qw_horner.cu) against the oracle."""

import numpy as np
import pytest

from horner_model import rk4_horner, step_coefficients
from oracle.rk4_numpy import rk4_L2

CASES = [
    (64, "circular", "lin", 40.0, 1.3, 400, "oracle"),
    (64, "circular", "cerf_3", 64.0, 0.8, 512, "oracle"),
    (32, "circular", "sqrt", 20.0, 2.0, 300, "oracle"),
    (40, "circular", "cbrt", 9.0, 0.0, 100, "oracle"),
    (64, "circular", "static", 30.0, 1.0, 500, "oracle"),
    (48, "circular", "cerf_7", 30.0, 1.7, 300, "laplacian"),
    (64, "complete", "lin", 20.0, 1.0 / 64, 800, "laplacian"),
    (64, "complete", "lin", 20.0, 0.0, 100, "laplacian"),       # a == 0: every ratio is 0/0
    (64, "complete", "cerf_3", 2.0, 40.0, 600, "oracle"),
    (16, "circular", "lin", 16.0, 1.0, 1024, "oracle"),         # BASELINE config 1
]


@pytest.mark.parametrize("n,topo,sched,T,g,S,go", CASES)
def test_nested_form_equals_rk4(n, topo, sched, T, g, S, go):
    p, norm, psi = rk4_horner(n, topo, sched, T, g, S, go)
    po, no, psio = rk4_L2(n, topo, sched, [T], [g], n_steps=S, gamma_on=go, return_psi=True)
    assert abs(p - po[0]) <= 1e-13 and abs(norm - no[0]) <= 1e-13
    assert np.abs(psi - psio[0]).max() <= 1e-13


@pytest.mark.parametrize("n,topo,sched,T,g,S,go", [c for c in CASES if c[1] == "circular"])
def test_addend_form_on_the_one_sided_window_equals_rk4(n, topo, sched, T, g, S, go):
    """What the kernels execute since round 2 (HornerStep in csrc/qw_horner_level.cuh): t_l =
    y_w + sigma_l as the addend of the marked slot, from the window y_w .. y_{w-3} only."""
    from horner_model import rk4_horner_addend

    for target in (None, 0, n - 1):
        p, norm, psi, asym = rk4_horner_addend(n, sched, T, g, S, go, target=target)
        po, no, psio = rk4_L2(n, topo, sched, [T], [g], n_steps=S, gamma_on=go, return_psi=True,
                              target=target)
        assert abs(p - po[0]) <= 1e-13 and abs(norm - no[0]) <= 1e-13
        assert np.abs(psi - psio[0]).max() <= 1e-13
        assert asym <= 1e-15            # numpy's roll-based stencil is symmetric to rounding


def test_static_schedule_gives_the_autonomous_ratios():
    """Equal operator scalars in all stages: r = h a (1, 1/2, 1/3, 1/4)."""
    rho, _ = step_coefficients(0.1, 0.7, 0.7, 0.7, -0.3, -0.3, -0.3, 2.0, 6.0)
    assert np.allclose(rho, 0.07 * np.array([1, 1 / 2, 1 / 3, 1 / 4]), rtol=1e-14)


def test_config1_value():
    p, _, _ = rk4_horner(16, "circular", "lin", 16.0, 1.0, 1024, "oracle")
    assert abs(p - 0.42038739054526) < 1e-12


COMPLETE_CASES = [
    (64, "lin", 20.0, 1.0 / 64, 800, "laplacian"),
    (51, "cerf_3", 30.0, 0.9 / 51, 1200, "laplacian"),
    (17, "cbrt", 0.8, 12.0, 300, "oracle"),
    (64, "sqrt", 2.0, 40.0, 600, "oracle"),
    (40, "static", 25.0, 1.3 / 40, 500, "laplacian"),
    (33, "lin", 400.0, 1.0 / 33, 20000, "laplacian"),      # long run: sums never re-anchored
]


@pytest.mark.parametrize("n,sched,T,g,S,go", COMPLETE_CASES)
def test_complete_graph_subsystem_map_equals_rk4(n, sched, T, g, S, go):
    """The scheme of rk4_complete_horner / rk4_complete_folded: tabulated linear map of
    (y_w, S(y)), level means, corrections; the tracked pair stays on the true one."""
    from horner_model import rk4_complete_mapped

    p, norm, psi, x = rk4_complete_mapped(n, sched, T, g, S, go)
    po, no, psio = rk4_L2(n, "complete", sched, [T], [g], n_steps=S, gamma_on=go, return_psi=True)
    assert abs(p - po[0]) <= 1e-12 and abs(norm - no[0]) <= 1e-12
    assert np.abs(psi - psio[0]).max() <= 1e-12
    w = (n - 1) // 2
    assert abs(x[0] - psio[0][w]) <= 1e-12 and abs(x[1] - psio[0].sum()) <= 1e-11


@pytest.mark.parametrize("n,sched,T,g,S,go", COMPLETE_CASES)
def test_complete_graph_tracked_stage_sums_equal_rk4(n, sched, T, g, S, go):
    """The scheme of rk4_complete_packed_any: S(stage input) from the marked vertex alone."""
    from horner_model import rk4_complete_tracked_sums

    p, norm, psi, sy = rk4_complete_tracked_sums(n, sched, T, g, S, go)
    po, no, psio = rk4_L2(n, "complete", sched, [T], [g], n_steps=S, gamma_on=go, return_psi=True)
    assert abs(p - po[0]) <= 1e-12 and abs(norm - no[0]) <= 1e-12
    assert np.abs(psi - psio[0]).max() <= 1e-12 and abs(sy - psio[0].sum()) <= 1e-11
