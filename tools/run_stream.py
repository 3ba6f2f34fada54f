#!/usr/bin/env python
"""Small streaming-kernel run for profiling: python tools/run_stream.py [kb] [points] [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

kb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
points = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
flag = {1: 8, 2: 16, 4: 24, 8: 0}[kb] | (256 if len(sys.argv) > 4 and sys.argv[4] == "r8" else 0)
dev = torch.device("cuda")
T = torch.linspace(0.1, 64.0, points, dtype=torch.float64, device=dev)
g = torch.linspace(0.0, 2.5, points, dtype=torch.float64, device=dev)
prob = qw.Problem(n=65536, topology="circular", schedule="lin", n_steps=steps, flags=flag)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    p, nrm = qw.rk4_batch(prob, T, g)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
print("kb=%d points=%d steps=%d: %.3f ms, %.3e site-updates/s, p[0]=%.12f"
      % (kb, points, steps, ms, points * 65536 * steps / (ms * 1e-3), float(p[0])))
