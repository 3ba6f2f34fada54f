#!/usr/bin/env python
"""A/B timing of kernel variants selected by qw_problem.flags (development tool).

    python tools/ab_kernels.py [--rows 16] [--reps 3] [--workload c5|c3|c4]

Prints site-updates/s per variant (CUDA events, best of reps) and checks that all variants
agree with the first to 1e-12.
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

VARIANTS = {
    "c5": [("horner R=16 occ6 (default)", 0), ("mirror-reduced, one warp per point", 8192), ("mirror-reduced, 8 warps/SM", 8192 | 128), ("horner R=16 occ4", 128), ("horner R=8", 256),
           ("horner R=16 occ6 no-rotate", 4096), ("horner R=16 occ4 no-rotate", 4096 | 128),
           ("horner R=8 no-rotate", 4096 | 256),
           ("v2 R=16", 2048), ("v2 R=8 occ3", 2048 | 256), ("v2 R=8 occ4", 2048 | 256 | 128),
           ("v2 R=8 mbarrier occ3", 256 | 64), ("v1 R=8 legacy occ3", 256 | 32),
           ("v1 R=8 legacy occ4", 256 | 32 | 128)],
    "c4": [("packed nested (default)", 0), ("mirror-reduced, four points per warp (drop-in default)", 8192),
           ("mirror-reduced, two points per warp", 8192 | 256), ("mirror-reduced, one warp per point", 8192 | 512),
           ("packed stage form", 2048), ("horner R=8 one warp per point", 512), ("v2 occ3", 512 | 2048), ("v2 occ4", 128), ("v1 legacy occ3", 32),
           ("v1 legacy occ4", 32 | 128)],
    "c2": [("packed (default)", 0), ("packed 16 warps/SM at 128 regs", 128), ("packed R<=8", 256), ("v2 R=2 CTA-per-point", 512)],
    "c3": [("complete horner R=16 (default)", 0), ("complete horner R=16, one aux warp", 128),
           ("small complete graph: block per point (no packing)", 512),
           ("complete horner R=8", 256),
           ("complete v2 R=16", 2048), ("complete barrier-free R=16", 1024),
           ("complete v2 R=8", 2048 | 256), ("complete v1 R=8 legacy", 256 | 32),
           ("complete R=8 exact-sum", 256 | 4)],
}
SHAPES = {
    "c5": dict(n=1024, topology="circular", schedule="lin", n_steps=4096, nT=1024, t=(0.1, 1024), g=(0, 2.5)),
    "c4": dict(n=256, topology="circular", schedule="lin", n_steps=512, nT=1000, t=(20, 30), g=(0.8, 1.2)),
    "c2": dict(n=64, topology="circular", schedule="lin", n_steps=1024, nT=100, t=(0.1, 64), g=(0, 2.5)),
    "c3": dict(n=4096, topology="complete", schedule="cerf_3", n_steps=2048, nT=256, t=(0.1, 512),
               g=(0.5 / 4096, 2.5 / 4096), gamma_on="laplacian"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--n", type=int, default=0, help="override the dimension of the shape")
    ap.add_argument("--only", default="", help="comma-separated flag values to run")
    a = ap.parse_args()
    sh = dict(SHAPES[a.workload])
    if a.n:
        sh["n"] = a.n
    dev = torch.device("cuda")
    t = torch.linspace(sh["t"][0], sh["t"][1], sh["nT"], dtype=torch.float64, device=dev)
    b = torch.linspace(sh["g"][0], sh["g"][1], a.rows, dtype=torch.float64, device=dev)
    ref = None
    only = [int(x) for x in a.only.split(",") if x]
    for name, flags in VARIANTS[a.workload]:
        if only and flags not in only:
            continue
        prob = qw.Problem(n=sh["n"], topology=sh["topology"], schedule=sh["schedule"],
                          n_steps=sh["n_steps"], gamma_on=sh.get("gamma_on", "oracle"),
                          flags=flags)
        pl = qw.plan(prob)
        best = 1e30
        for _ in range(a.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p, nrm = qw.rk4_heatmap(prob, t, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if ref is None:
            ref = p.clone()
        err = float((p - ref).abs().max())
        rate = a.rows * sh["nT"] * sh["n"] * sh["n_steps"] / (best * 1e-3)
        print("%-22s %8.2f ms  %.4e site-updates/s  regs=%d blocks/SM=%d  |dp|max vs first=%.1e"
              % (name, best, rate, pl["regs_per_thread"], pl["blocks_per_sm"], err))


if __name__ == "__main__":
    main()
