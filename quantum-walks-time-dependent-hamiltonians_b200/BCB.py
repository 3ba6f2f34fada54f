"""Drop-in for the hot path of the reference's ``BCB.py``: both topologies, string
schedules, the time-independent comparator and the parallel heatmap.

Module globals as in BCB.py:20-24 (+ ``type_of_qw``): ``dimension``, ``step_function``
('lin', 'sqrt', 'cbrt', 'cerf_3', 'cerf_5', 'cerf_7'), ``topology`` ('circular',
'complete'), ``type_of_qw`` ('dependent', 'independent').  ``parallel_routine`` writes the
same three .npy files (BCB.py:298-311); under ``torchrun`` it shards the beta rows over
the GPUs instead of over Ray workers.
"""

from . import _dropin

dimension = None
step_function = "lin"
topology = "circular"
type_of_qw = "dependent"
rtolerance = 1e-6
atolerance = 1e-6
n_steps = None
step_size = 1e-3
gamma_on = "oracle"
integrator = "rk4"           # 'rk45': SciPy's adaptive RK45 as the reference ships it
ti_method = "exact"         # time-independent comparator: 'exact' (Chebyshev) or 'rk4'
device = None
mirror = True                # symmetry-reduced state (QW_FLAG_MIRROR), exact for the uniform start every
                             # caller uses (BCB.py:179-180): half the ring / (psi_w, psi_other) on the
                             # complete graph, same p, norm and psi; False integrates every site

_dropin.install(globals(), "BCB")
