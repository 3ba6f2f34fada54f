#!/usr/bin/env python
"""Where does the thesis' production sweep (BCB_latest.py:309-320) spend its time on the GPU?
Times the 140 heatmaps (75 beta x 4 T, step_size 1e-3) without the file output: all on 8 streams,
and a few dimensions alone (latency of one heatmap = latency of its longest column)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402
from qwtdh_b200 import BCB_latest as bl  # noqa: E402

dev = torch.device("cuda:0")
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
beta = torch.tensor(bl.PRODUCTION_BETA, dtype=torch.float64, device=dev)


def tarr(n):
    return torch.tensor([(np.pi / 4) * np.sqrt(n), np.sqrt(n), (np.pi / 2) * np.sqrt(n),
                         2 * np.sqrt(n)], dtype=torch.float64, device=dev)


def sweep(streams):
    pool = [torch.cuda.Stream(device=dev) for _ in range(streams)]
    times = {n: tarr(n) for n in range(3, 72, 2)}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 0
    for step in ("lin", "sqrt", "cbrt", "cerf_3"):
        for n in range(3, 72, 2):
            prob = qw.Problem(n=n, topology="circular", schedule=step, step_size=1e-3, flags=flags)
            with torch.cuda.stream(pool[k % streams]):
                qw.rk4_heatmap(prob, times[n], beta)
            k += 1
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return t1 - t0, time.perf_counter() - t0


def fused():
    probs, ts = [], []
    for step in ("lin", "sqrt", "cbrt", "cerf_3"):
        for n in range(3, 72, 2):
            probs.append(qw.Problem(n=n, topology="circular", schedule=step, step_size=1e-3,
                                    flags=flags))
            ts.append(tarr(n))
    torch.cuda.synchronize()
    l0 = qw.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    qw.rk4_heatmaps(probs, ts, [beta] * len(probs))
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return t1 - t0, time.perf_counter() - t0, e0.elapsed_time(e1), qw.launch_count() - l0


fused()
issue, total, gpu_ms, launches = fused()
print("fused (qw_rk4_multi_grid_dev): issue %.1f ms  total %.1f ms  GPU %.1f ms  %d launches for 140 "
      "heatmaps" % (issue * 1e3, total * 1e3, gpu_ms, launches))
for streams in (1, 16):
    sweep(streams)
    issue, total = sweep(streams)
    print("streams=%2d  issue %.1f ms  total %.1f ms" % (streams, issue * 1e3, total * 1e3))
for n in (3, 7, 15, 16, 31, 33, 51, 63, 64, 65, 71):
    prob = qw.Problem(n=n, topology="circular", schedule="lin", step_size=1e-3, flags=flags)
    t = tarr(n)
    qw.rk4_heatmap(prob, t, beta)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    qw.rk4_heatmap(prob, t, beta)
    e1.record()
    torch.cuda.synchronize()
    S = int(2 * np.sqrt(n) / 1e-3)
    pl = qw.plan(prob)
    print("N=%3d  %-18s R=%d pts/warp=%d  %.3f ms  %d steps  %.0f ns/step" % (
        n, pl["path"], pl["sites_per_thread"], pl["steps_per_pass"], e0.elapsed_time(e1), S,
        e0.elapsed_time(e1) * 1e6 / S))
