"""Drop-in for the class API of the reference's ``QW_Search.py``.

``Hamiltonian(dim, type, topology, target, step_func, gamma)`` and ``State(dim)`` keep
their constructor and method names (QW_Search.py:58-65, 124-147).  The reference's
``State.evolve`` only prints ``_runge_kutta()`` for the time-dependent type and applies an
element-wise ``np.exp`` for the time-independent one (QW_Search.py:140-147); here both
run the fixed-step RK4 kernel with the stub's rule ``step_size = 1e-3``,
``n = int(time / step_size)`` (QW_Search.py:153-154).

Encodings (QW_Search.py:70-81, 104-109, 142-147): topology 0 complete / 1 cycle;
step_func 0 cerf_3 / 1 lin / 2 sqrt; type 0 time-independent / 1 time-dependent.
"""

import numpy as np

from . import _dropin, engine

step_size = 1e-3           # QW_Search.py:153
n_steps = None             # set to use h = time / n_steps instead
device = None
mirror = True              # exact symmetry reduction for the uniform start (QW_FLAG_MIRROR)

_TOPOLOGY = {0: "complete", 1: "circular"}


class Hamiltonian:
    """Hamiltonian class"""

    def __init__(self, dim, type, topology, target, step_func, gamma):
        self.dim = int(dim)
        self.type = int(type)
        self.topology = int(topology)
        self.target = int(target)
        self.step_func = int(step_func)
        self.gamma = float(gamma)
        self.L = None
        self.H = None

    def build_laplacian(self):
        if self.topology not in _TOPOLOGY:
            raise ValueError("undefined topology: %r" % (self.topology,))
        self.L = _dropin.laplacian(self.dim, _TOPOLOGY[self.topology])

    def build_hamiltonian(self):
        """Time-independent H = L - gamma|target><target|; the time-dependent one starts
        as H = L (QW_Search.py:88-97, without the aliasing that corrupts L there)."""
        if self.L is None:
            self.build_laplacian()
        self.H = self.L.copy()
        if self.type == 0:
            self.H[self.target, self.target] -= self.gamma

    def update_hamiltonian(self, t_time, T_time):
        tau = float(t_time) / T_time
        g_T = _dropin.schedule_value(tau, {0: "cerf_3", 1: "lin", 2: "sqrt"}[self.step_func])
        self.H = (1 - g_T) * self.L
        self.H[self.target, self.target] += -g_T * self.gamma

    def exp_H(self):
        return np.exp(self.H)

    def problem(self):
        sched = "static" if self.type == 0 else self.step_func
        rule = dict(n_steps=n_steps) if n_steps is not None else dict(step_size=step_size)
        return engine.Problem(n=self.dim, topology=self.topology, schedule=sched,
                              target=self.target, flags=8192 if mirror else 0, **rule)


class State:
    """Quantum state class"""

    def __init__(self, dim):
        self.dim = int(dim)
        self.ket = None
        self.H = None
        self._elapsed = 0.0       # time already evolved since set_initial (static H only)

    def set_initial(self, type=0):
        if type == 0:
            self.ket = np.ones((self.dim), dtype=np.complex128) / np.sqrt(self.dim)
            self._elapsed = 0.0
        else:
            print("Initial type missing or not yet implemented")

    def set_hamiltonian(self, hamiltonian):
        self.H = hamiltonian

    def bra(self):
        return np.transpose(np.conjugate(self.ket))

    def evolve(self, time):
        """Advance ``ket`` by ``time`` with the RK4 kernel.  The kernels start from the uniform
        state (the only initial state the reference defines, QW_Search.py:128-131), so a
        repeated call re-integrates from it over the accumulated time: well defined for the
        time-independent type (exp(-iH t2) exp(-iH t1) = exp(-iH (t1 + t2))); for the
        time-dependent type the schedule g(t / T) belongs to ONE run of length T and a second
        ``evolve`` raises."""
        if self.H.type not in (0, 1):
            print("Error, unknown evolution type")
            return
        if self._elapsed > 0.0:
            if self.H.type == 1:
                raise ValueError("time-dependent evolution is one run of length T from the "
                                 "uniform state: call set_initial(0) before evolving again")
            total = self._elapsed + float(time)
            self.ket = _runge_kutta(self.H, None, total)
            self._elapsed = total
            return
        self.ket = _runge_kutta(self.H, self.ket, time)
        self._elapsed = float(time)

    def success_probability(self):
        """|<target|psi>|^2 / <psi|psi>."""
        k = self.ket
        return float(abs(k[self.H.target]) ** 2 / np.vdot(k, k).real)


def _runge_kutta(state, psi, time):
    """Evolves state using 4th order Runge-Kutta integrator (QW_Search.py:149-164).
    ``state`` is the Hamiltonian; ``psi`` must be the uniform state the kernels start
    from (the only initial state the reference defines, QW_Search.py:128-131)."""
    uniform = np.ones(state.dim) / np.sqrt(state.dim)
    if psi is not None and not np.allclose(psi, uniform, rtol=0, atol=1e-15):
        raise ValueError("only the uniform initial state (set_initial(0)) is supported")
    _, _, out = engine.rk4_batch(state.problem(), [float(time)], [state.gamma],
                                 return_psi=True, device=device)
    return out[0]


if __name__ == '__main__':
    HC = Hamiltonian(5, 0, 0, 1, 0, 10)
    HC.build_laplacian()
    HC.build_hamiltonian()
    psi = State(5)
    psi.set_initial(0)
    psi.set_hamiltonian(HC)
    print(psi.ket)
    psi.evolve(10)
    print(psi.ket)
