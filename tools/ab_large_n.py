#!/usr/bin/env python
"""Throughput above 4096 sites: site-updates/s of what the library selects, next to the
global-memory catch-all (QW_FLAG_FORCE_GENERIC = 1), the streaming kernel (QW_FLAG_FORCE_STREAM = 2)
and the symmetry reductions (QW_FLAG_MIRROR = 8192).

    python tools/ab_large_n.py [--topology complete|circular] [--sizes 8192,16384] [--steps 512]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--topology", default="complete")
ap.add_argument("--sizes", default="4096,5000,8192,12289,16384")
ap.add_argument("--steps", type=int, default=512)
ap.add_argument("--flags", default="0,1")  # 128: cluster kernel with one big CTA per SM
ap.add_argument("--points", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda")
for n in [int(x) for x in a.sizes.split(",")]:
    pts = a.points or max(296, int(2 ** 31 // (n * a.steps)) // 296 * 296)
    cplt = a.topology == "complete"
    T = torch.linspace(0.1, 50.0 if not cplt else 300.0, pts, dtype=torch.float64, device=dev)
    g = torch.linspace(0.2, 2.5, pts, dtype=torch.float64, device=dev) / (n if cplt else 1.0)
    ref = None
    for flags in [int(x) for x in a.flags.split(",")]:
        prob = qw.Problem(n=n, topology=a.topology, schedule="cerf_3", n_steps=a.steps,
                          gamma_on="laplacian" if cplt else "oracle", flags=flags)
        pl = qw.plan(prob)
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p, nrm = qw.rk4_batch(prob, T, g)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        if ref is None:
            ref = p.clone()
        print("%s N=%-6d flags=%-5d %-16s R=%-2d thr=%-3d blk/SM=%d x%-2d %5d pts %8.2f ms  %.4e "
              "site-updates/s  |dp| vs first %.1e" % (
                  a.topology, n, flags, pl["path"], pl["sites_per_thread"], pl["threads_per_block"],
                  pl["blocks_per_sm"], pl["steps_per_pass"], pts, best,
                  pts * n * a.steps / (best * 1e-3), float((p - ref).abs().max())), flush=True)
