#!/usr/bin/env python
"""C4-shaped run for profiling: cycle N=256, 1e5 points, n_steps=512 (warp-packed kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda")
T = torch.linspace(20.0, 30.0, 100000, dtype=torch.float64, device=dev)
g = torch.linspace(0.8, 1.2, 100000, dtype=torch.float64, device=dev)
prob = qw.Problem(n=256, topology="circular", schedule="lin", n_steps=512, flags=flags)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    p, nrm = qw.rk4_batch(prob, T, g)
    e1.record()
    torch.cuda.synchronize()
print("%s: %.2f ms, %.3e site-updates/s" % (qw.plan(prob)["path"], e0.elapsed_time(e1),
                                           1e5 * 256 * 512 / (e0.elapsed_time(e1) * 1e-3)))
