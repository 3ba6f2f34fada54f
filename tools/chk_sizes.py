import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import qwtdh_b200 as qw
from oracle import c_oracle
T, g = np.array([3.0, 11.0, 40.0, 300.0]), np.array([0.4, 1.1, 2.3, 2.5])
for n, fl in [(1024, 0), (1024, 8192), (512, 0), (512, 8192), (2048, 0), (2048, 8192), (256, 0), (3000, 0), (300, 8192)]:
    prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=2048, flags=fl)
    p, nrm = qw.rk4_batch(prob, T, g)
    po, no, _ = c_oracle.rk4_batch(n, "circular", "lin", T, g, n_steps=2048)
    print(n, fl, qw.plan(prob)["path"], np.abs(p - po).max(), np.abs(nrm - no).max())
