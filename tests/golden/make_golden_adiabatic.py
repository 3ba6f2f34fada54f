#!/usr/bin/env python
"""Golden values of the reference's adiabatic diagnostics (Schroedinger_Solver.py:146-280),
produced by running the reference's own functions in the build container:

    python tests/golden/make_golden_adiabatic.py        # writes golden_adiabatic_v1.json

compute_energy_diff(s, beta) and compute_gamma(s, beta) on a set of (dimension,
step_function, s, beta), and adiabatic_theorem_check([_, _, beta]) for the linear schedule
(for sqrt / cbrt the coupling diverges at s -> 0 and the reference's shgo result is not a
well-defined number).  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

warnings.filterwarnings("ignore")


def main():
    out = {"source": "Schroedinger_Solver.py compute_energy_diff / compute_gamma / "
                     "adiabatic_theorem_check, numpy %s" % np.__version__,
           "points": [], "checks": []}
    for dim, sched in [(15, 1), (15, 2), (15, 3), (21, 1), (21, 2), (51, 1), (51, 3), (64, 1)]:
        ref = ref_loader.load("Schroedinger_Solver", dimension=dim, step_function=sched)
        for beta in (0.35, 1.02212, 1.9):
            for s in (0.02, 0.11, 0.3, 0.5, 0.64, 0.77, 0.9, 0.98):
                gap = float(ref.compute_energy_diff(s, beta))
                gam = float(np.ravel(ref.compute_gamma(s, beta))[0])       # = -|<phi1|dH|phi0>|
                out["points"].append({"dimension": dim, "step_function": sched, "s": s,
                                      "beta": beta, "gap": gap, "gamma": -gam})
    for dim, beta in [(15, 1.02212), (21, 0.9), (31, 1.2)]:
        ref = ref_loader.load("Schroedinger_Solver", dimension=dim, step_function=1)
        res = ref.adiabatic_theorem_check([0.0, 0.0, beta])
        out["checks"].append({"dimension": dim, "step_function": 1, "beta": beta,
                              "adiabatic_time": float(res[0]), "gamma_max": float(res[1]),
                              "energy_min": float(res[2]), "flag": res[3]})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_adiabatic_v1.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote %s: %d points, %d checks" % (path, len(out["points"]), len(out["checks"])))
    for c in out["checks"]:
        print(c)


if __name__ == "__main__":
    main()
