#!/usr/bin/env python
"""The production sweep as ONE fused call (for `ncu --metrics gpu__time_duration.sum`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402
from qwtdh_b200 import BCB_latest as bl  # noqa: E402

dev = torch.device("cuda:0")
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
beta = torch.tensor(bl.PRODUCTION_BETA, dtype=torch.float64, device=dev)
probs, ts = [], []
for step in ("lin", "sqrt", "cbrt", "cerf_3"):
    for n in range(3, 72, 2):
        probs.append(qw.Problem(n=n, topology="circular", schedule=step, step_size=1e-3, flags=flags))
        ts.append(torch.tensor([(np.pi / 4) * np.sqrt(n), np.sqrt(n), (np.pi / 2) * np.sqrt(n),
                                2 * np.sqrt(n)], dtype=torch.float64, device=dev))
for _ in range(2):
    qw.rk4_heatmaps(probs, ts, [beta] * len(probs))
torch.cuda.synchronize()
