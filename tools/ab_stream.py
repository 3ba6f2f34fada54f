#!/usr/bin/env python
"""A/B of the streaming kernel on the C5 streaming shape (cycle N=65536, 32x32 grid, S=256):
steps per pass 1 / 2 / 4 / 8; bulk-copy (TMA) data movement at 4 / 5 (16 sites per thread) or
3 / 4 (8 sites) blocks per SM against the cp.async kernel (QW_FLAG_CYCLE_LEGACY = 32), the stage
form (QW_FLAG_NO_HORNER = 2048) and the half-ring reduction (QW_FLAG_MIRROR = 8192).

    python tools/ab_stream.py [--n 65536] [--quick]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--steps", type=int, default=256)
ap.add_argument("--quick", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda")
n, S = a.n, a.steps
side = 32 if n >= 65536 else int(round((1024 * 65536 / n) ** 0.5))
t = torch.linspace(0.1, 64.0, side, dtype=torch.float64, device=dev)
b = torch.linspace(0.0, 2.5, side, dtype=torch.float64, device=dev)
ref = {}
variants = [("tma", 0), ("tma+1blk", 128), ("cp.async", 32)]
if not a.quick:
    variants += [("stage", 2048), ("mirror", 8192)]
for kb, kflag in ((1, 8), (2, 16), (4, 24), (8, 0)):
    for name, hflag in variants:
        for r8 in (0, 256):
            if kb == 8 and r8:
                continue
            if a.quick and r8 and kb != 1:
                continue
            prob = qw.Problem(n=n, topology="circular", schedule="lin", n_steps=S,
                              flags=2 | kflag | hflag | r8)
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
            rate = side * side * n * S / (best * 1e-3)
            print("KB=%d %-9s R=%-2d %7.2f ms  %.4e site-updates/s  %.0f GB/s algorithmic  regs=%d "
                  "blocks/SM=%d smem=%d |dp| vs first %.1e" % (
                      kb, name, pl["sites_per_thread"], best, rate, 32 * rate / 1e9,
                      pl["regs_per_thread"], pl["blocks_per_sm"], pl["smem_bytes"],
                      float((p - ref["p"]).abs().max())), flush=True)
