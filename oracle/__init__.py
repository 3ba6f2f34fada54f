"""CPU oracle for the RK4 heatmap hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and there only as the checker (or the
timed CPU baseline) -- never as the thing shipped.  The product package
``quantum-walks-time-dependent-hamiltonians_b200`` never imports it.

Tiers:

* ``ref_loader``   loads the reference's own functions, unmodified, from /root/reference
                   (only available in the build container; used to make ``tests/golden``).
* ``rk4_numpy.rk4_L0``  classic RK4 time loop over the reference's *own*
                   ``schrodinger_equation`` (BCB.py:163-174).
* ``rk4_numpy.rk4_L1``  same loop, dense H(t) restated after BCB.py:28-99.
* ``rk4_numpy.rk4_L2``  implicit stencil / rank-one restatement, batched over points.
* ``rk4_oracle.c``      the same restatement in plain C + OpenMP (large sizes and the timed
                   host baseline).

Parity pin: the reference has no tests and no golden vectors, and it
never implemented the fixed-step RK4 it sketches (QW_Search.py:149-164).  The oracle is
therefore pinned against outputs of the reference's own functions run in the build
container: H(t) matrices, RHS vectors, RK4-over-reference-RHS trajectories, RK45 and expm
probabilities, all stored in ``tests/golden/`` by ``tests/golden/make_golden.py``.
``gamma_on='laplacian'`` has no reference code at all: *parity unpinned* beyond the thesis'
analytic points (Eq. 1.28, Eq. 2.14) and the two-variable symmetric-subspace ODE.
"""

from .defs import (  # noqa: F401
    TOPOLOGY_COMPLETE, TOPOLOGY_CYCLE, SCHEDULES, SCHEDULE_NAMES, GAMMA_ON_ORACLE,
    GAMMA_ON_LAPLACIAN, schedule_id, topology_id, gamma_on_id, default_target,
)
