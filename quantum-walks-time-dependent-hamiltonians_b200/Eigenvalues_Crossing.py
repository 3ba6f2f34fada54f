"""Drop-in for the reference's ``Eigenvalues_Crossing.py``: the spectrum of
H(t) = (1 - t/T) L - (t/T) gamma |w><w| on the cycle graph as a function of t, the data behind
the level-crossing figure.

The reference builds the dense matrices and calls ``scipy.linalg.eig`` once per sample
(Eigenvalues_Crossing.py:17-71); here ``compute_eigenvalues`` and the sampling loop of the
``__main__`` block (``eigenvalues_sweep``) evaluate every root of the secular equation of the
rank-one perturbation on the device, one launch for all samples.  The dense builders are kept
for inspection with the reference's names and results (its integer division
``(dimension-1)/2`` is Python 2; the marked vertex is ``(dimension-1)//2`` as everywhere else).
"""

import numpy as np

from . import engine

schedule = "lin"             # the reference's g = time / T (Eigenvalues_Crossing.py:63)
topology = "circular"
device = None


def generate_loop_hamiltonian(dimension):
    """Laplacian of the ring, diag 2 / neighbours -1 (Eigenvalues_Crossing.py:17-47)."""
    n = int(dimension)
    lap = 2.0 * np.eye(n)
    idx = np.arange(n)
    lap[idx, (idx + 1) % n] -= 1.0
    lap[idx, (idx - 1) % n] -= 1.0
    return lap


def generate_oracle_hamiltonian(dimension):
    """|w><w| with w = (dimension - 1) // 2 (Eigenvalues_Crossing.py:49-56)."""
    n = int(dimension)
    ham = np.zeros((n, n))
    ham[(n - 1) // 2, (n - 1) // 2] = 1.0
    return ham


def time_dependent_hamiltonian(dimension, gamma, time, T):
    """(1 - g) L - g gamma |w><w|, g = time / T (Eigenvalues_Crossing.py:58-63)."""
    g = float(time) / T
    return ((1 - g) * generate_loop_hamiltonian(dimension)
            - g * gamma * generate_oracle_hamiltonian(dimension))


def compute_eigenvalues(dimension, gamma, time, TIME):
    """Sorted eigenvalues of H(time) (Eigenvalues_Crossing.py:65-71) -> float64[dimension]."""
    prob = engine.Problem(n=int(dimension), topology=topology, schedule=schedule)
    return engine.spectrum_batch(prob, [float(time) / TIME], [float(gamma)], device=device)[0]


def eigenvalues_sweep(dimension=15, gamma=1.02212, T=10.0, sampling=100, filename=None):
    """The sampling loop of the reference's ``__main__`` (Eigenvalues_Crossing.py:88-112):
    ``sampling`` times t_i = i * T^2 / 100, i = 1 .. sampling, of a walk of total time T^2;
    returns eigenvalues[sampling][dimension] and, if ``filename`` is given, writes it as the
    reference does (``'15_quadratic.txt'``, format ``%.3e``)."""
    total = float(T) * float(T)
    time_step = total / 100
    times = time_step * np.arange(1, int(sampling) + 1)
    prob = engine.Problem(n=int(dimension), topology=topology, schedule=schedule)
    eig = engine.spectrum_batch(prob, times / total, np.full(times.size, float(gamma)),
                                device=device)
    if filename:
        np.savetxt(filename, eig, fmt='%.3e')
    return eig
