#!/usr/bin/env python
"""A/B of the streaming kernel on the C5 streaming shape (cycle N=65536, 32x32 grid, S=256):
steps per pass 1 / 2 / 4 / 8, nested form vs stage form (QW_FLAG_NO_HORNER = 2048)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

dev = torch.device("cuda")
n, S = 65536, 256
t = torch.linspace(0.1, 64.0, 32, dtype=torch.float64, device=dev)
b = torch.linspace(0.0, 2.5, 32, dtype=torch.float64, device=dev)
ref = {}
for kb, kflag in ((1, 8), (2, 16), (4, 24), (8, 0)):
    for name, hflag in (("nested", 0), ("stage", 2048), ("mirror", 8192)):
        for r8 in (0, 256):
            if kb == 8 and r8:
                continue
            prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S,
                              flags=kflag | hflag | r8)
            pl = qw.plan(prob)
            best = 1e30
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                p, nrm = qw.rk4_heatmap(prob, t, b)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ref.setdefault("p", p.clone())
            rate = 1024 * n * S / (best * 1e-3)
            print("KB=%d %-6s R=%-2d %7.2f ms  %.4e site-updates/s  %.0f GB/s algorithmic  regs=%d "
                  "blocks/SM=%d  |dp| vs first %.1e" % (kb, name, pl["sites_per_thread"], best, rate,
                                                       32 * rate / 1e9, pl["regs_per_thread"],
                                                       pl["blocks_per_sm"],
                                                       float((p - ref["p"]).abs().max())))
