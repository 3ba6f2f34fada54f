#!/usr/bin/env python
"""Generate tests/golden/golden_v1.{json,npz} by running the REFERENCE's own functions.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests, so these fixtures are the pin for the oracle:
every number below is produced by the reference's unmodified code loaded through
oracle/ref_loader.py (H(t), the RHS, RK45 probabilities, expm probabilities, the
robustness formula), or by a classic RK4 time loop whose right-hand side is the reference's
unmodified ``schrodinger_equation`` (tier L0).  numpy/scipy versions are recorded.
"""

import contextlib
import io
import json
import os
import sys

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, rk4_numpy as rk  # noqa: E402

OUT_DIR = os.path.dirname(os.path.abspath(__file__))
SS_SCHEDULE = {"lin": 1, "sqrt": 2, "cbrt": 3}  # Schroedinger_Solver.py:55-60


def oracle_state(n):
    w = np.zeros((n, 1))
    w[int((n - 1) / 2)][0] = 1
    return w


def main():
    arrays, meta = {}, {"numpy": np.__version__, "scipy": scipy.__version__,
                        "hamiltonians": [], "rhs": [], "rk4_L0": [], "rk45": [],
                        "static": [], "heatmaps": [], "robustness": []}
    rng = np.random.default_rng(20200912)

    bcb = ref_loader.load("BCB", rtolerance=1e-6, atolerance=1e-6,
                          type_of_qw="dependent")
    ss = ref_loader.load("Schroedinger_Solver", rtolerance=1e-6, atolerance=1e-6)
    rob = ref_loader.load("Robustness")

    # 1. dense H(t) from generate_hamiltonian (BCB.py:28-99, SS:21-70)
    k = 0
    for (n, topo, sched, beta, t, T) in [
            (5, "circular", "lin", 1.0, 0.0, 4.0), (8, "circular", "lin", 0.7, 1.3, 4.0),
            (9, "circular", "sqrt", 2.0, 2.5, 5.0), (7, "circular", "cbrt", 1.5, 0.4, 3.0),
            (8, "circular", "cerf_3", 1.1, 2.9, 3.0), (6, "circular", "cerf_5", 0.3, 1.0, 7.0),
            (11, "circular", "cerf_7", 2.5, 6.5, 7.0), (3, "circular", "lin", 1.0, 0.5, 1.0),
            (5, "complete", "lin", 3.0, 0.8, 2.0), (12, "complete", "cerf_3", 10.0, 1.9, 2.0),
            (4, "complete", "sqrt", 0.5, 0.0, 1.0), (6, "complete", "lin", 2.0, 1.0, 0.0)]:
        bcb.dimension, bcb.topology, bcb.step_function = n, topo, sched
        H = np.array(bcb.generate_hamiltonian(n, beta, t, T), dtype=np.float64)
        arrays["H_%d" % k] = H
        meta["hamiltonians"].append(dict(key="H_%d" % k, src="BCB", n=n, topology=topo,
                                         schedule=sched, beta=beta, t=t, T=T))
        k += 1
    for (n, sched, beta, t, T) in [(5, "lin", 1.0, 2.0, 4.0), (10, "sqrt", 0.9, 1.0, 9.0),
                                   (7, "cbrt", 2.2, 3.0, 3.0)]:
        ss.dimension, ss.step_function = n, SS_SCHEDULE[sched]
        H = np.array(ss.generate_hamiltonian(n, beta, t, T), dtype=np.float64)
        arrays["H_%d" % k] = H
        meta["hamiltonians"].append(dict(key="H_%d" % k, src="SS", n=n,
                                         topology="circular", schedule=sched, beta=beta,
                                         t=t, T=T))
        k += 1

    # 2. RHS vectors from schrodinger_equation (BCB.py:163-174) on seeded random states
    for k, (n, topo, sched, beta, t, T) in enumerate([
            (16, "circular", "lin", 1.0, 3.0, 16.0), (33, "circular", "cerf_3", 1.7, 5.5, 9.0),
            (64, "circular", "sqrt", 0.4, 10.0, 64.0), (15, "complete", "cbrt", 3.0, 0.6, 2.0),
            (40, "complete", "lin", 25.0, 0.01, 0.3)]):
        bcb.dimension, bcb.topology, bcb.step_function = n, topo, sched
        y = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        dy = np.asarray(bcb.schrodinger_equation(t, y, beta, T), dtype=complex)
        arrays["rhs_y_%d" % k] = y
        arrays["rhs_dy_%d" % k] = dy
        meta["rhs"].append(dict(key=k, n=n, topology=topo, schedule=sched, beta=beta,
                                t=t, T=T))

    # 3. RK4 over the reference's own RHS (tier L0), both step rules
    cases = [
        (16, "circular", "lin", 16.0, 1.0, dict(n_steps=256)),
        (16, "circular", "lin", 16.0, 1.0, dict(n_steps=1024)),
        (16, "circular", "cerf_3", 16.0, 1.3, dict(n_steps=1024)),
        (16, "circular", "sqrt", 10.0, 0.7, dict(n_steps=1024)),
        (17, "circular", "cbrt", 12.0, 2.0, dict(n_steps=1024)),
        (15, "circular", "cerf_5", 9.0, 1.9, dict(n_steps=512)),
        (21, "circular", "cerf_7", 11.0, 0.9, dict(n_steps=512)),
        (15, "complete", "lin", 2.0, 3.0, dict(n_steps=1024)),
        (15, "complete", "cbrt", 2.0, 3.0, dict(n_steps=1024)),
        (32, "complete", "cerf_3", 1.0, 20.0, dict(n_steps=512)),
        (5, "circular", "lin", 7.0, 1.5, dict(step_size=1e-2)),
        (9, "circular", "sqrt", 3.337, 0.8, dict(step_size=1e-2)),
        (3, "circular", "lin", 2.0, 1.0, dict(n_steps=128)),
        (16, "circular", "lin", 0.0, 1.0, dict(n_steps=64)),
        (51, "circular", "lin", 51.0, 1.2, dict(n_steps=512)),
        (64, "circular", "lin", 20.0, 1.0, dict(n_steps=256)),
    ]
    for k, (n, topo, sched, T, beta, rule) in enumerate(cases):
        bcb.dimension, bcb.topology, bcb.step_function = n, topo, sched
        p, norm, psi = rk.rk4_L0(bcb, T, beta, **rule)
        arrays["L0_psi_%d" % k] = psi
        meta["rk4_L0"].append(dict(key=k, n=n, topology=topo, schedule=sched, T=T,
                                   beta=beta, rule=rule, p=p, norm=norm))
        print("L0", n, topo, sched, T, beta, rule, p, norm)

    # 4. the reference as shipped: RK45, tol 1e-6 (BCB.py:193-203) -- loose cross-check
    for (n, topo, sched, T, beta) in [
            (16, "circular", "lin", 16.0, 1.0), (16, "circular", "cerf_3", 16.0, 1.3),
            (16, "circular", "sqrt", 10.0, 0.7), (17, "circular", "cbrt", 12.0, 2.0),
            (15, "complete", "lin", 2.0, 3.0), (15, "complete", "cbrt", 2.0, 3.0),
            (51, "circular", "lin", 51.0, 1.2), (64, "circular", "cerf_3", 8.0, 2.5)]:
        bcb.dimension, bcb.topology, bcb.step_function = n, topo, sched
        val = bcb.evaluate_probability([beta, T], oracle_state(n))
        assert val.shape == (1,)
        meta["rk45"].append(dict(n=n, topology=topo, schedule=sched, T=T, beta=beta,
                                 p=float(-val[0])))

    # 5. time-independent comparator: expm (BCB.py:137-159) and RK4 over the dense TI H
    for (n, t, gamma, S) in [(5, 3.0, 1.0, 512), (51, 20.0, 0.85, 2048),
                             (64, 32.0, 0.55, 4096), (16, 6.0, 1.4, 1024)]:
        bcb.dimension, bcb.topology = n, "circular"
        val = bcb.evaluate_time_independent_probability(t, gamma, oracle_state(n))
        assert val.shape == (1, 1)
        p, norm, _ = rk.rk4_L0_static(bcb, t, gamma, n_steps=S)
        meta["static"].append(dict(n=n, time=t, gamma=gamma, n_steps=S,
                                   p_expm=float(-val[0, 0]), p_rk4=p, norm_rk4=norm))

    # 6. heatmaps, layout [beta, T]: L0 on a small grid; SS.grid_probability_evaluation
    n, topo, sched = 8, "circular", "lin"
    bcb.dimension, bcb.topology, bcb.step_function = n, topo, sched
    tarr, barr = np.linspace(0.1, 8, 4), np.linspace(0, 2.5, 3)
    grid = np.empty((3, 4))
    for i, T in enumerate(tarr):
        for j, b in enumerate(barr):
            grid[j][i] = rk.rk4_L0(bcb, T, b, n_steps=128)[0]
    arrays["heat_L0_0"] = grid
    meta["heatmaps"].append(dict(key="heat_L0_0", kind="L0", n=n, topology=topo,
                                 schedule=sched, time=[0.1, 8, 4], beta=[0, 2.5, 3],
                                 n_steps=128))
    ss.dimension, ss.step_function = 5, 1
    ss.heatmap2d = lambda *a, **k: None       # plotting is out of scope (matplotlib absent)
    with contextlib.redirect_stdout(io.StringIO()):
        g45 = ss.grid_probability_evaluation(5, 1, 20, 4, 0.1, 5, 3, 0.92, 7)
    arrays["heat_RK45_0"] = np.asarray(g45, dtype=np.float64)
    meta["heatmaps"].append(dict(key="heat_RK45_0", kind="RK45", n=5, topology="circular",
                                 schedule="lin", time=[1, 20, 4], beta=[0.1, 5, 3]))

    # 7. Robustness.py:25-31 on a seeded synthetic heatmap
    prob = rng.random((12, 5))
    arrays["rob_prob"] = prob
    for (bi, ti, v) in [(5, 2, 2), (5, 2, 4), (6, 0, 1), (2, 4, 2)]:
        meta["robustness"].append(dict(beta_index=bi, time_index=ti, variation=v,
                                       R=rob.robustness(prob, bi, ti, v)))

    np.savez_compressed(os.path.join(OUT_DIR, "golden_v1.npz"), **arrays)
    with open(os.path.join(OUT_DIR, "golden_v1.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("wrote", len(arrays), "arrays")


if __name__ == "__main__":
    main()
