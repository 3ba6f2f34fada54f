#!/usr/bin/env python
"""BASELINE config 2 (cycle N=64, 100 x 100 heatmap, n_steps=1024): lin, cerf_3 and static as
three launches against ONE multi-problem launch, full ring and half ring (QW_FLAG_MIRROR)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qwtdh_b200 as qw  # noqa: E402

dev = torch.device("cuda")
t = torch.linspace(0.1, 64.0, 100, dtype=torch.float64, device=dev)
b = torch.linspace(0.0, 2.5, 100, dtype=torch.float64, device=dev)
peak = qw.fp64_peak_tflops(300.0)


def timed(fn, reps=5):
    fn()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for flags, name in ((0, "full ring"), (8192, "half ring (QW_FLAG_MIRROR)")):
    probs = [qw.Problem(n=64, topology="circular", schedule=s, n_steps=1024, flags=flags)
             for s in ("lin", "cerf_3", "static")]
    su = 3 * 10000 * 64 * 1024
    one = timed(lambda: qw.rk4_heatmap(probs[0], t, b))
    seq = timed(lambda: [qw.rk4_heatmap(p, t, b) for p in probs])
    fus = timed(lambda: qw.rk4_heatmaps(probs, [t] * 3, [b] * 3))
    print("%s: %s" % (name, qw.plan(probs[0])))
    for label, ms, work in (("one schedule, one launch", one, su / 3), ("three launches", seq, su),
                            ("one multi-problem launch", fus, su)):
        rate = work / (ms * 1e-3)
        print("  %-26s %7.3f ms  %.3e full-ring site-updates/s = %.1f %% of the FP64 roofline "
              "(52 flops, peak %.1f TFLOP/s)" % (label, ms, rate, 100 * 52 * rate / 1e12 / peak, peak))
