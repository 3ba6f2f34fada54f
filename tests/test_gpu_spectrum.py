"""Full spectrum of H(s) on the device (csrc/qw_adiabatic.cu, spectrum_kernel) against dense
diagonalisation (oracle.rk4_numpy.spectrum_dense, restating Eigenvalues_Crossing.py:59-71), and
the drop-in module ``Eigenvalues_Crossing``."""

import os

import numpy as np
import pytest

from oracle import rk4_numpy as rk

pytestmark = pytest.mark.gpu

TOL = 2e-12            # absolute, on eigenvalues of size <= 4 + |d|


@pytest.mark.parametrize("n", [3, 4, 5, 15, 16, 63, 64, 255, 256, 1000])
def test_cycle_spectrum_vs_dense(qw, n):
    rng = np.random.default_rng(n)
    for sched, gamma_on in (("lin", "oracle"), ("cbrt", "oracle"), ("cerf_3", "laplacian"),
                            ("sqrt", "laplacian")):
        s = np.concatenate([[0.0, 1.0, 0.5, 1e-9, 1 - 1e-9], rng.uniform(0.0, 1.0, 11)])
        g = np.concatenate([[1.0, 1.0, 0.0, 2.0, 2.0], rng.uniform(0.0, 3.0, 11)])
        prob = qw.Problem(n=n, topology="circular", schedule=sched, gamma_on=gamma_on)
        eig = qw.spectrum_batch(prob, s, g)
        assert eig.shape == (s.size, n)
        for q in range(s.size):
            ref = rk.spectrum_dense(n, "circular", sched, s[q], g[q], gamma_on=gamma_on)
            assert np.abs(eig[q] - ref).max() <= TOL, (n, sched, gamma_on, s[q], g[q])
            assert np.all(np.diff(eig[q]) >= -1e-15)               # ascending without a sort


def test_cycle_spectrum_negative_gamma_on_oracle(qw):
    """gamma < 0 with the oracle placement makes the marked vertex repulsive (d > 0): the roots
    move up instead of down and the interlacing order flips."""
    n = 17
    s, g = np.array([0.3, 0.7, 0.999]), np.array([-1.5, -0.2, -4.0])
    eig = qw.spectrum_batch(qw.Problem(n=n, topology="circular", schedule="lin"), s, g)
    for q in range(3):
        ref = rk.spectrum_dense(n, "circular", "lin", s[q], g[q])
        assert np.abs(eig[q] - ref).max() <= TOL


@pytest.mark.parametrize("n", [1, 2, 3, 16, 1000])
def test_complete_spectrum_vs_dense(qw, n):
    rng = np.random.default_rng(n)
    for gamma_on, scale in (("oracle", float(n)), ("laplacian", 1.0 / n)):
        s = np.concatenate([[0.0, 1.0], rng.uniform(0.0, 1.0, 8)])
        g = np.concatenate([[1.0, 1.0], rng.uniform(0.0, 2.0, 8)]) * scale
        prob = qw.Problem(n=n, topology="complete", schedule="lin", gamma_on=gamma_on)
        eig = qw.spectrum_batch(prob, s, g)
        for q in range(s.size):
            ref = rk.spectrum_dense(n, "complete", "lin", s[q], g[q], gamma_on=gamma_on)
            assert np.abs(eig[q] - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_spectrum_consistent_with_adiabatic_gap(qw):
    n = 101
    s = np.linspace(0.05, 0.95, 19)
    prob = qw.Problem(n=n, topology="circular", schedule="lin")
    eig = qw.spectrum_batch(prob, s, np.full(s.size, 1.3))
    gap, _, e0 = qw.adiabatic_batch(prob, s, np.full(s.size, 1.3))
    assert np.abs(eig[:, 0] - e0).max() <= 1e-13
    assert np.abs((eig[:, 1] - eig[:, 0]) - gap).max() <= 1e-13


def test_dimension_limit_is_reported(qw):
    with pytest.raises(ValueError, match="spectrum"):
        qw.spectrum_batch(qw.Problem(n=5000, topology="circular", schedule="lin"), [0.5], [1.0])


def test_eigenvalues_crossing_dropin(qw, tmp_path):
    """The module's functions with the reference's names, and the sampling loop of its __main__
    (dimension 15, gamma 1.02212, T 10, 100 samples, '15_quadratic.txt')."""
    from qwtdh_b200 import Eigenvalues_Crossing as ec

    H = ec.time_dependent_hamiltonian(15, 1.02212, 30.0, 100.0)
    assert H.shape == (15, 15) and H[7, 7] == pytest.approx(0.7 * 2 - 0.3 * 1.02212)
    assert H[0, 14] == pytest.approx(-0.7) and H[0, 1] == pytest.approx(-0.7)
    ev = ec.compute_eigenvalues(15, 1.02212, 30.0, 100.0)
    assert np.abs(ev - np.linalg.eigvalsh(H)).max() <= TOL
    out = tmp_path / "15_quadratic.txt"
    eig = ec.eigenvalues_sweep(filename=str(out))
    assert eig.shape == (100, 15)
    for i in (0, 49, 99):                                  # t_i = (i + 1) T^2 / 100
        Hi = ec.time_dependent_hamiltonian(15, 1.02212, (i + 1) * 1.0, 100.0)
        assert np.abs(eig[i] - np.linalg.eigvalsh(Hi)).max() <= TOL
    back = np.loadtxt(str(out))
    assert back.shape == (100, 15) and np.abs(back - eig).max() <= 5e-3 * (1 + np.abs(eig).max())
    assert os.path.getsize(str(out)) > 0
