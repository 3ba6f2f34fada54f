"""B200-native fixed-step RK4 sweeps for quantum-walk search with time-dependent
Hamiltonians -- a drop-in for the hot path of
mgarbellini/Quantum-Walks-Time-Dependent-Hamiltonians.

    engine                  batched / heatmap / sharded calls over the C ABI
    Schroedinger_Solver, BCB, BCB_latest, Runge_Kutta_Error, QW_Search, Robustness,
    heatmap_generator       modules with the reference's own call signatures and globals

The directory name carries hyphens (it mirrors the upstream repository name); import it
as ``qwtdh_b200`` (alias module at the repository root) or through ``importlib``.
"""

from . import _lib, engine                                     # noqa: F401
from .engine import (Problem, rk4_batch, rk4_heatmap, rk4_heatmaps, rk4_batch_host,    # noqa: F401
                     rk4_heatmap_host, rhs, rk45_batch, rk45_heatmap, ti_exact_batch, ti_exact_heatmap, ensemble_stats, robustness_map, heatmap_summary, maximize_probability, plan,
                     adiabatic_batch, adiabatic_check, spectrum_batch,
                     fp64_peak_tflops, launch_count, shard_rows, heatmap_sharded, sharded_d2h_bytes,
                     longest_first_order)

__version__ = "0.1.0"
