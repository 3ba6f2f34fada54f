"""Drop-in for the reference's ``BCB_latest.py``: BCB with the thesis' production sweep
(75 beta values x 4 times {pi/4, 1, pi/2, 2} * sqrt(N); BCB_latest.py:242-264) and the
file names ``Robustness.py`` reads (BCB_latest.py:287-289)."""

import numpy as np

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

# 0.1 ... 3.95 in steps of 0.05 without 1.05, 2.05, 3.05 (BCB_latest.py:242-245)
PRODUCTION_BETA = [round(0.05 * k, 2) for k in range(2, 80) if k not in (21, 41, 61)]

_dropin.install(globals(), "BCBL")


def production_sweep(step_functions=("lin", "sqrt", "cbrt", "cerf_3"),
                     dims=tuple(np.arange(3, 71 + 1, 2)), streams=16, fused=True):
    """The __main__ loop of BCB_latest.py:309-320.

    A production heatmap is 300 points on a ring of at most 71 sites: one launch occupies a few
    per cent of the GPU and runs for as long as its longest point.  ``fused`` (default; one
    process, fixed-step RK4, time-dependent walk): all heatmaps of the sweep go to the library in
    ONE call (``engine.rk4_heatmaps`` -> ``qw_rk4_multi_grid_dev``), which lays the problems that
    share a warp-packed kernel end to end in one grid -- a handful of launches instead of 140;
    the files written are bitwise those of the sequential loop.  ``fused=False`` and
    ``streams`` > 1 issue one launch per heatmap round-robin on that many CUDA streams (round 1).
    ``streams`` = 1 with ``fused=False``, a process group, 'rk45' or the time-independent
    comparator take the reference's sequential loop."""
    global step_function, dimension
    import time as _time

    import torch
    import torch.distributed as dist

    from . import engine

    sequential = (integrator != "rk4" or type_of_qw != "dependent"
                  or (dist.is_available() and dist.is_initialized())
                  or (not fused and streams <= 1))
    if sequential:
        for step in step_functions:
            step_function = step
            for dim in dims:
                dimension = int(dim)
                parallel_routine()                         # noqa: F821 (installed above)
        return
    dev = engine._device(device)
    beta = np.asarray(PRODUCTION_BETA, dtype=np.float64)
    d_beta = torch.from_numpy(beta).to(dev)
    tic = _time.perf_counter()
    work = []
    for step in step_functions:
        for dim in dims:
            n = int(dim)
            time_array = [(np.pi / 4) * np.sqrt(n), np.sqrt(n), (np.pi / 2) * np.sqrt(n),
                          2 * np.sqrt(n)]
            prob = engine.Problem(n=n, topology=topology, schedule=step, gamma_on=gamma_on,
                                  n_steps=n_steps, step_size=None if n_steps else step_size,
                                  flags=(8192 if mirror else 0) | _dropin.FIXED_GEOMETRY)
            # (fixed geometry, as parallel_routine: the files do not depend on how the sweep is issued)
            work.append((n, step, time_array, prob,
                         torch.tensor(time_array, dtype=torch.float64, device=dev)))
    if fused:
        with torch.cuda.device(dev):
            outs = engine.rk4_heatmaps([w[3] for w in work], [w[4] for w in work],
                                       [d_beta] * len(work), device=dev)
        pending = [(n, step, time_array, p) for (n, step, time_array, _, _), (p, _) in
                   zip(work, outs)]
    else:
        pool = [torch.cuda.Stream(device=dev) for _ in range(int(streams))]
        pending = []
        for n, step, time_array, prob, d_time in work:
            st = pool[len(pending) % len(pool)]
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                p, _ = engine.rk4_heatmap(prob, d_time, d_beta, device=dev)
            pending.append((n, step, time_array, p))
    torch.cuda.synchronize(dev)
    for n, step, time_array, p in pending:
        np.save(str(n) + '_probability_' + step + '.npy', p.cpu().numpy())
        np.save(str(n) + '_circ_time.npy', time_array)
        np.save(str(n) + '_circ_beta.npy', list(PRODUCTION_BETA))
        print('Success: N ', n, ' in ', int((_time.perf_counter() - tic) / 60), ' min')
